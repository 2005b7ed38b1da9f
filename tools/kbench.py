#!/usr/bin/env python
"""Per-kernel timing of the pass kernels (CUDA events on the launching stream) for the BASELINE
workloads.  Used while tuning; the judged numbers come from bench.py."""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import B_ALG, workload_model  # noqa: E402
from classicalspinmc.jl_b200 import _lib  # noqa: E402


def _unit_cell(name):
    from classicalspinmc.jl_b200 import workloads
    uc = workloads.unit_cell(name)
    if not uc.basis:
        uc.basis.append(np.zeros(uc.D))
    return uc


def time_cycles(eng, stream, n, orc, mc):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    eng.cycles_async(max(2, n // 10), orc, mc)
    torch.cuda.synchronize()
    e0.record(stream)
    eng.cycles_async(n, orc, mc)
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e-3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workloads", default="C2,C2:4096,C3:256:8,C4:32:16,C5:512:1")
    ap.add_argument("--n", type=int, default=200)
    ap.add_argument("--flags", type=int, default=0)
    ap.add_argument("--ssf", type=int, default=0, help="also time the structure factor at this many wavevectors")
    args = ap.parse_args()
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    peak = 6547.8
    md_uc = {}
    for spec in args.workloads.split(","):
        parts = spec.split(":")
        name = parts[0]
        L = int(parts[1]) if len(parts) > 1 else None
        R = int(parts[2]) if len(parts) > 2 else 1
        md, cfg = workload_model(name, L)
        md_uc[spec] = _unit_cell(name)
        eng = _lib.Engine(md, n_replicas=R, seed=1, stream=stream.cuda_stream, flags=args.flags)
        eng.randomize(7)
        eng.set_temperatures(np.geomspace(0.5, 2.0, R))
        N, C = eng.N, eng.n_colours
        balg = B_ALG.get(C, 24.0 * (C + 1))
        flops_upd = eng.kernel_costs()[0]
        fp64_peak = torch.cuda.get_device_properties(0).multi_processor_count * 64 * 2 * 1.965e9 / 1e12   # nominal
        out = {"workload": spec, "N": N, "R": R, "colours": C, "mode": eng.kernel_mode, "groups": eng.sweep_groups(),
               "persist": eng.persist_info(), "blocks": eng.replica_blocks()[0], "flops_per_or_update": flops_upd}
        for label, orc, mc in (("or", 1, 0), ("or2", 2, 0), ("or10", 10, 0), ("or40", 40, 0), ("metro", 0, 1), ("metro2", 0, 2), ("metro10", 0, 10), ("cycle10+1", 10, 1)):
            n = max(args.n // max(orc + mc, 1), 5)
            dt = time_cycles(eng, stream, n, orc, mc)
            upd = n * (orc + mc) * N * R
            out[label] = {"Gupd_s": upd / dt / 1e9, "us_per_pass": dt / (n * (orc + mc) * C) * 1e6,
                          "GBs_alg": upd * balg / dt / 1e9, "frac": upd * balg / dt / 1e9 / peak}
            if mc == 0:
                out[label]["fp64_tflops"] = upd * flops_upd / dt / 1e12
                out[label]["fp64_frac_nominal"] = upd * flops_upd / dt / 1e12 / fp64_peak
        if args.ssf:
            # equal-time structure factor: n_k wavevectors x N sites (src/spin_correlations.jl:6-43)
            import time
            uc = md_uc[spec]
            ks = np.random.default_rng(1).uniform(-np.pi, np.pi, size=(uc.D, args.ssf))
            eng.structure_factor(uc.lattice_vectors, uc.basis, ks)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(5):
                eng.structure_factor(uc.lattice_vectors, uc.basis, ks)
            dt = (time.perf_counter() - t0) / 5
            out["ssf"] = {"n_k": args.ssf, "ms": dt * 1e3, "Gpairs_s": args.ssf * N / dt / 1e9}
        print(json.dumps(out))
        eng.close()


if __name__ == "__main__":
    main()
