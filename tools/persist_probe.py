#!/usr/bin/env python
"""Runs sequences of overrelaxation / Metropolis sweeps of one workload on the tile-resident kernel (or, with
--no-persist, the pass kernels) — the command ncu wraps for the profiles of csmc_persist, and a quick A/B timer."""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from classicalspinmc.jl_b200 import _abi, _lib, workloads  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="C2")
    ap.add_argument("--L", type=int, default=None)
    ap.add_argument("--replicas", type=int, default=1)
    ap.add_argument("--or-sweeps", type=int, default=20)
    ap.add_argument("--metro-sweeps", type=int, default=0)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--no-persist", action="store_true")
    args = ap.parse_args()
    md, _ = workloads.workload_model(args.workload, args.L)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    flags = _abi.FLAG_NO_AUTOTUNE | (_abi.FLAG_NO_PERSIST if args.no_persist else _abi.FLAG_PERSIST)
    eng = _lib.Engine(md, n_replicas=args.replicas, seed=1, stream=stream.cuda_stream, flags=flags)
    eng.randomize(7)
    eng.set_temperatures(np.geomspace(0.5, 2.0, args.replicas))
    out = {"workload": args.workload, "persist": eng.persist_info()}
    for _ in range(2):
        eng.cycles_async(1, args.or_sweeps, args.metro_sweeps)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    eng.cycles_async(args.reps, args.or_sweeps, args.metro_sweeps)
    e1.record(stream)
    torch.cuda.synchronize()
    n_pass = args.reps * (args.or_sweeps + args.metro_sweeps) * eng.n_colours
    out["us_per_pass"] = e0.elapsed_time(e1) * 1e3 / n_pass
    out["Gupd_s"] = args.reps * (args.or_sweeps + args.metro_sweeps) * eng.N * args.replicas / (e0.elapsed_time(e1) * 1e-3) / 1e9
    print(json.dumps(out))
    eng.close()


if __name__ == "__main__":
    main()
