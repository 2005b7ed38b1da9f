#!/usr/bin/env python
"""One (10 OR + 1 Metropolis) cycle of a workload between cudaProfilerStart / Stop, after warm-up: the command ncu
wraps to see the colour passes INSIDE the sequencing the library actually uses (replica blocks, time-skewed strips,
replica groups) instead of one isolated, cold-cache launch.

    ncu --profile-from-start off --cache-control none --clock-control none \
        --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \
        -k regex:csmc_sweep --csv --log-file gpurun_out/x.csv python tools/ncu_cycle.py --workload C3 --replicas 64

With --cache-control none and a metric set that fits one pass, each launch is measured once in the cache state the
preceding (unprofiled or profiled) launches left behind, so dram__bytes shows what the blocks / strips save.
tools/ncu_cycle.py --summarise x.csv prints per-kernel totals."""
import argparse
import collections
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def summarise(path):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    i_name, i_metric, i_val, i_id = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
    i_grid = hdr.index("Grid Size")
    launches = collections.OrderedDict()
    for r in rows[1:]:
        d = launches.setdefault(r[i_id], {"kernel": r[i_name], "grid": r[i_grid]})
        d[r[i_metric]] = float(r[i_val].replace(",", ""))
    tot = collections.defaultdict(lambda: collections.defaultdict(float))
    for d in launches.values():
        t = tot[d["kernel"]]
        t["launches"] += 1
        t["us"] += d.get("gpu__time_duration.sum", 0.0) / 1e3
        t["dram_read_MB"] += d.get("dram__bytes_read.sum", 0.0) / 1e6
        t["dram_write_MB"] += d.get("dram__bytes_write.sum", 0.0) / 1e6
    print(f"# {path}: {len(launches)} launches in the profiled cycle")
    for k, t in tot.items():
        print(f"{k:28s} launches {int(t['launches']):5d}  time {t['us']:9.1f} us  dram read {t['dram_read_MB']:9.1f} MB  write {t['dram_write_MB']:9.1f} MB")
    first = list(launches.values())[:12]
    print("# first launches (kernel, grid, us, dram read MB, dram write MB):")
    for d in first:
        print(f"  {d['kernel']:24s} {d['grid']:>16s} {d.get('gpu__time_duration.sum', 0) / 1e3:8.1f} {d.get('dram__bytes_read.sum', 0) / 1e6:9.2f} {d.get('dram__bytes_write.sum', 0) / 1e6:9.2f}")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--summarise", default=None)
    ap.add_argument("--workload", default="C2")
    ap.add_argument("--L", type=int, default=None)
    ap.add_argument("--replicas", type=int, default=1)
    ap.add_argument("--cycles", type=int, default=1)
    ap.add_argument("--flags", type=int, default=0)
    args = ap.parse_args()
    if args.summarise:
        return summarise(args.summarise)
    import numpy as np
    import torch

    from classicalspinmc.jl_b200 import _lib, workloads
    md, _ = workloads.workload_model(args.workload, args.L)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    eng = _lib.Engine(md, n_replicas=args.replicas, seed=1, stream=stream.cuda_stream, flags=args.flags)
    eng.randomize(7)
    eng.set_temperatures(np.geomspace(0.5, 2.0, args.replicas))
    eng.cycles_async(3, 10, 1)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    eng.cycles_async(args.cycles, 10, 1)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print(json.dumps({"workload": args.workload, "L": args.L, "replicas": args.replicas, "blocks": eng.replica_blocks()[0],
                      "groups": eng.sweep_groups()[0], "skew": eng.skew_info(), "persist": eng.persist_info()[:3]}))
    eng.close()


if __name__ == "__main__":
    main()
