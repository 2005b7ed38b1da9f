#!/usr/bin/env python
"""Finite-temperature Monte Carlo of a pyrochlore magnet by parallel tempering (the reference's
examples/parallel_tempering/runner.jl on the GPU engine).

The reference runs one temperature per MPI rank; here one process drives one GPU and holds a block of
temperatures:

    python examples/parallel_tempering/runner.py OUT/ --temperatures 128                   # one GPU
    torchrun --nproc-per-node 8 examples/parallel_tempering/runner.py OUT/ --temperatures 128   # 8 GPUs, 16 each

One configuration_<slot>.h5 per temperature slot is written, plus IC_<slot>/ measurement snapshots for the
slots listed with --save-ic.
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import classicalspinmc.jl_b200 as csm  # noqa: E402
import input_file as inp  # noqa: E402
from pyrochlore import addInteractionsLocal  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("outpath", nargs="?", default=os.getcwd() + "/")
ap.add_argument("--temperatures", type=int, default=None, help="total number of temperatures (default: 16 per process)")
ap.add_argument("--L", type=int, default=inp.L)
ap.add_argument("--t-thermalization", type=int, default=inp.t_thermalization)
ap.add_argument("--t-measurement", type=int, default=inp.t_measurement)
ap.add_argument("--B", type=float, default=0.0, help="field strength in tesla")
ap.add_argument("--save-ic", type=int, nargs="*", default=[0], help="temperature slots whose measurement snapshots are kept")
args = ap.parse_args()

# one process per GPU (torchrun); a single process otherwise
world, rank = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0))
if world > 1:
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
    dist.init_process_group("nccl")

# logarithmically spaced temperatures, split into contiguous blocks (the first processes take the remainder)
n_T = args.temperatures or 16 * world
temps = np.geomspace(inp.Tmin, inp.Tmax, n_T)
per, rem = divmod(n_T, world)
lo = rank * per + min(rank, rem)
T = temps[lo:lo + per + (1 if rank < rem else 0)]

P = csm.Pyrochlore()
addInteractionsLocal(P, {"Jxx": inp.Jxx, "Jyy": inp.Jyy, "Jzz": inp.Jzz})
for b, hb in enumerate(inp.h_local, start=1):
    csm.addZeemanCoupling(P, b, hb * args.B * inp.mu_B)

lat = csm.Lattice((args.L,) * 3, P, inp.S)
params = {"t_thermalization": args.t_thermalization, "t_measurement": args.t_measurement,
          "probe_rate": inp.probe_rate, "swap_rate": inp.swap_rate, "overrelaxation_rate": inp.overrelaxation,
          "report_interval": inp.report_interval, "checkpoint_rate": inp.checkpoint_rate}
device = int(os.environ.get("LOCAL_RANK", 0))
mc = csm.MonteCarlo(T, lat, params, outpath=args.outpath, device=device)

csm.parallel_tempering(mc, args.save_ic)

if world > 1:
    dist.destroy_process_group()
