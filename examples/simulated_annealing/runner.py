#!/usr/bin/env python
"""Ground state of the Kitaev honeycomb model in a field by simulated annealing + deterministic updates
(the reference's examples/simulated_annealing/runner.jl on the GPU engine).

    python examples/simulated_annealing/runner.py [output directory]
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import classicalspinmc.jl_b200 as csm  # noqa: E402
from honeycomb import addInteractionsKitaev  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("outpath", nargs="?", default=os.getcwd() + "/")
ap.add_argument("--L", type=int, default=4)
ap.add_argument("--t-thermalization", type=int, default=int(1e4))
ap.add_argument("--t-deterministic", type=int, default=int(1e6))
args = ap.parse_args()

# lattice and interaction parameters
L, S = args.L, 1.0
K, h = -1.0, 0.1
h_vec = h * np.array([-1.0, 1.0, 0.0]) / np.sqrt(2)
inparams = {"K": K, "h": h}                       # human-readable inputs, stored as attributes of the params file

# Monte Carlo parameters and target temperature
mcparams = {"t_thermalization": args.t_thermalization, "t_deterministic": args.t_deterministic, "overrelaxation_rate": 10}
T = 1e-7

H = csm.Honeycomb()
addInteractionsKitaev(H, {"K": K})
csm.addZeemanCoupling(H, 1, h_vec)
csm.addZeemanCoupling(H, 2, h_vec)

lat = csm.Lattice((L, L), H, S)
mc = csm.MonteCarlo(T, lat, mcparams, outpath=args.outpath, outprefix="configuration", inparams=inparams)

csm.simulated_annealing(mc, lambda x: 1.0 * 0.9 ** x, 1.0)
csm.deterministic_updates(mc)
csm.write_MC_checkpoint(mc)
print("energy per site:", csm.energy_density(mc.lattice))
