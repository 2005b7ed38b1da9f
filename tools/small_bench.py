#!/usr/bin/env python
"""Small-lattice regime (the sizes the reference's examples actually run): parallel tempering on the
pyrochlore L=8 example (N=2048, 128 temperatures, swap every 50, 10 OR : 1 Metropolis) and annealing of
the L=4 README lattice.  Compares the resident kernel and the per-colour pass kernels; the CPU figure next to
them is bench.py's reference-arm leg (bench.cpu_pt_baseline), the one place outside tests/ that runs oracle/."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from classicalspinmc.jl_b200 import _lib  # noqa: E402
from classicalspinmc.jl_b200._abi import FLAG_JIT, FLAG_NO_RESIDENT, ModelData  # noqa: E402
import bench  # noqa: E402
from classicalspinmc.jl_b200 import workloads as models  # noqa: E402


def pt_case(md, R, sweeps, flags):
    eng = _lib.Engine(md, n_replicas=R, seed=1, flags=flags)
    eng.randomize(3)
    eng.pt_init(np.geomspace(0.09 / 11.6, 14 / 11.6, R))
    p = dict(t_thermalization=10 ** 9, t_measurement=0, probe_rate=2000, swap_rate=50, overrelaxation_rate=10)
    eng.pt_run(p, 0, 100)
    l0 = eng.launches
    t0 = time.perf_counter()
    eng.pt_run(p, 100, 100 + sweeps)
    dt = time.perf_counter() - t0
    upd = sweeps * 1.1 * eng.N * R
    return {"mode": eng.kernel_mode, "sweeps_per_s": sweeps / dt, "Gupd_s": upd / dt / 1e9, "launches": eng.launches - l0}


def main():
    md = ModelData(models.pyrochlore_local(), (8, 8, 8), 0.5)
    R = 128
    out = {"workload": "pyrochlore L=8 (N=2048), PT 128 temperatures"}
    out["resident"] = pt_case(md, R, 5000, FLAG_JIT)
    out["pass_kernels"] = pt_case(md, R, 1000, FLAG_JIT | FLAG_NO_RESIDENT)
    threads = os.cpu_count() or 1
    v, _, _ = bench.cpu_pt_baseline(md, np.geomspace(0.09 / 11.6, 14 / 11.6, R), 200, threads)
    out["cpu_reference_arm"] = {"threads": threads, "sweeps_per_s": v / (1.1 * 2048 * R), "Gupd_s": v / 1e9}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
