#!/usr/bin/env python
"""Short single-GPU check of the time-skewed strips (CSMC_FLAG_SKEW): bit-identity with the pass-by-pass order on
two small lattices (1 MiB budget so that the strips are used), then the HBM-bound point C2 at L=4096 (384 MiB):
throughput of the 10 OR + 1 Metropolis cycle pass by pass and strip by strip, and identity of the results.
No torch import.  One JSON line per item."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from classicalspinmc.jl_b200 import _lib  # noqa: E402
from classicalspinmc.jl_b200._abi import (FLAG_JIT, FLAG_NO_AUTOTUNE, FLAG_NO_GRAPH, FLAG_NO_RESIDENT, FLAG_SKEW,  # noqa: E402
                                          ModelData)
from classicalspinmc.jl_b200 import workloads as models  # noqa: E402


def out(**kw):
    print(json.dumps(kw), flush=True)


def small(name, builder, shape, bc="periodic", extra=0):
    os.environ["CSMC_L2_BLOCK_MB"] = "1"
    os.environ["CSMC_SWEEP_GROUPS"] = "1"
    md = ModelData(builder(), shape, 1.0, bc=bc)
    R, T, res, info = 2, np.array([0.7, 1.3]), [], {}
    for flags in (FLAG_JIT | FLAG_NO_RESIDENT | extra, FLAG_JIT | FLAG_NO_RESIDENT | FLAG_SKEW | extra):
        os.environ["CSMC_SKEW"] = "1" if flags & FLAG_SKEW else "0"
        eng = _lib.Engine(md, n_replicas=R, seed=77, flags=flags)
        usable, rows, reach, budget = eng.skew_info()
        eng.randomize(5)
        eng.set_temperatures(T)
        l0 = eng.launches
        eng.cycles_async(2, 2, 1)
        eng.sync()
        dl = eng.launches - l0
        eng.overrelax(4)
        eng.deterministic(2)
        res.append((np.stack([eng.get_spins(r) for r in range(R)]), eng.accepted().copy()))
        if flags & FLAG_SKEW:
            plan = _lib.skew_schedule(rows, 3 * eng.n_colours, reach, budget)
            info = dict(usable=usable, rows=rows, reach=reach, budget=budget, plan_launches=len(plan), cycle_launches=int(dl),
                        strips_used=bool(len(plan) > 0 and dl == 2 * R * len(plan) + (0 if extra & FLAG_NO_GRAPH else 2)))
        eng.close()
    same = bool(np.array_equal(res[0][0], res[1][0]) and np.array_equal(res[0][1], res[1][1]))
    out(check="skew_identical", case=name, identical=same, accepted=res[0][1].tolist(), **info)
    os.environ.pop("CSMC_L2_BLOCK_MB")
    os.environ.pop("CSMC_SWEEP_GROUPS")


def big(L, n_cycles=10):
    md = ModelData(models.square_heisenberg(J=-1.0, h=(0.0, 0.0, 0.1)), (L, L), 1.0)
    sig = []
    for flags in (0, FLAG_SKEW):
        t0 = time.perf_counter()
        os.environ["CSMC_SKEW"] = "1" if flags & FLAG_SKEW else "0"
        eng = _lib.Engine(md, n_replicas=1, seed=3, flags=flags)
        t_create = time.perf_counter() - t0
        eng.randomize(7)
        eng.set_temperatures(1.0)
        eng.cycles_async(2, 10, 1)
        eng.sync()
        best = 1e30
        for _ in range(3):
            t0 = time.perf_counter()
            eng.cycles_async(n_cycles, 10, 1)
            eng.sync()
            best = min(best, time.perf_counter() - t0)
        E = float(eng.total_energy()[0])
        acc = float(eng.accepted()[0])
        M = [float(v).hex() for v in eng.magnetization_vector()[0]]
        out(check="skew_timing", L=L, skew=bool(flags), Gupd_s=n_cycles * 11.0 * eng.N / best / 1e9, E=E.hex(), accepted=acc, M=M,
            skew_info=eng.skew_info(), create_s=round(t_create, 2), launches=int(eng.launches))
        sig.append((E, acc, M))
        eng.close()
    out(check="skew_identical_big", L=L, identical=bool(sig[0] == sig[1]))


def main():
    t0 = time.perf_counter()
    if "more" in sys.argv[1:]:
        small("square-open-512x128", models.square_heisenberg, (512, 128), bc="open", extra=FLAG_NO_AUTOTUNE)
        small("square-256-plain-launches", models.square_heisenberg, (256, 256), extra=FLAG_NO_GRAPH)
        small("triangular-multispin-1024x64", models.triangular_multispin, (1024, 64), extra=FLAG_NO_AUTOTUNE)
        out(t=round(time.perf_counter() - t0, 1))
        big(8192, 5)
        out(t=round(time.perf_counter() - t0, 1))
        return
    small("square-256", models.square_heisenberg, (256, 256))
    small("honeycomb-J3-512x64", lambda: models.kitaev_honeycomb(J3=0.25), (512, 64))
    out(t=round(time.perf_counter() - t0, 1))
    big(4096)
    out(t=round(time.perf_counter() - t0, 1))


if __name__ == "__main__":
    main()
