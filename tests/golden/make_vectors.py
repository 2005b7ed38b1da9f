#!/usr/bin/env python
"""Generates the committed golden fixtures of tests/golden/ (run from the repository root):

    python tests/golden/make_vectors.py

* reference_known_answers.json — every golden value / known answer the reference's own tests hold for the hot
  path (the reference is pure Julia and cannot run in the build image, so these are transcribed with their
  file:line; tests/test_oracle_golden.py pins the oracle against them);
* oracle_vectors.npz — seeded inputs and the oracle's outputs (local fields, site energies, total energy,
  magnetisation, overrelaxation / deterministic / same-stream Metropolis sweeps in colour order) for six small
  lattices covering every term kind and both boundary conditions.  The GPU parity tests compare the CUDA path
  with these files directly, so that check does not depend on the oracle library being rebuilt identically
  on the GPU box.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from classicalspinmc.jl_b200 import _lib  # noqa: E402
from classicalspinmc.jl_b200._abi import ModelData  # noqa: E402
from oracle import oracle as orc  # noqa: E402
from tests import models  # noqa: E402

CASES = {
    "square-8x8": (lambda: models.square_heisenberg(), (8, 8), "periodic", 1.0),
    "honeycomb-J3-6x4": (lambda: models.kitaev_honeycomb(J3=0.25), (6, 4), "periodic", 1.0),
    "pyrochlore-3x2x4": (lambda: models.pyrochlore_local(), (3, 2, 4), "periodic", 0.5),
    "triangular-multispin-onsite-8x4": (lambda: models.triangular_multispin(onsite=np.diag([0.1, -0.2, 0.3])), (8, 4), "periodic", 1.0),
    "mixed-basis-open-5x3": (lambda: models.mixed_basis_multispin(), (5, 3), "open", 0.8),
    "chain-open-17": (lambda: models.chain_heisenberg(), (17,), "open", 1.0),
}
SEED, T = 424242, 0.7

KNOWN = [
    {"cite": "test/latticetests.jl:3-7", "what": "norm of every random initial spin, rounded to 9 digits", "value": 1.0},
    {"cite": "test/latticetests.jl:10-17", "what": "total_energy, 1x1 square lattice, Zeeman h=(1,0,0), spin (1,0,0)", "value": -1.0},
    {"cite": "test/latticetests.jl:18", "what": "get_local_field of that site", "value": [-1.0, -0.0, -0.0]},
    {"cite": "test/latticetests.jl:21-30", "what": "total_energy / size, 2x2 square ferromagnet J=-I, spins along x", "value": -2.0},
    {"cite": "test/mctests.jl:36-49", "what": "round(E/N, 4) after simulated_annealing! + deterministic_updates!, Kitaev-Gamma-Gamma' honeycomb L=4 in a [111] field, Metropolis()", "value": -0.6444},
    {"cite": "test/mctests.jl:52-58", "what": "same with MetropolisAdaptive()", "value": -0.6444},
]


def build(name):
    builder, shape, bc, S = CASES[name]
    md = ModelData(builder(), shape, S, bc)
    lat = orc.OracleLattice(md)
    colour = _lib.plan(md)[0]
    order = (np.argsort(colour, kind="stable") + 1).astype(np.int64)      # Engine.colour_order()
    s0 = lat.randomize(seed=33)
    out = {"spins0": s0, "order": order, "field": lat.local_field_all(s0), "site_energy": lat.site_energy_all(s0)}
    E, E_abs = lat.total_energy(s0, with_abs=True)
    out["energy"] = np.array([E, E_abs])
    out["magnetization"] = lat.magnetization(s0, vector=True)
    s = s0.copy(); lat.overrelax(s, order, 3); out["or3"] = s
    s = s0.copy(); lat.deterministic(s, order, 2); out["det2"] = s
    s = s0.copy()
    acc = sum(lat.metropolis_philox(s, order, T, SEED, 0, sweep) for sweep in range(2))
    out["metro2"] = s
    out["metro2_accepted"] = np.array([acc])
    return out


def main():
    data = {}
    for name in CASES:
        for k, v in build(name).items():
            data[f"{name}/{k}"] = v
    np.savez_compressed(os.path.join(HERE, "oracle_vectors.npz"), **data)
    with open(os.path.join(HERE, "reference_known_answers.json"), "w") as f:
        json.dump({"source": "emilyzinnia/ClassicalSpinMC.jl test/ (transcribed; Julia cannot run here)", "answers": KNOWN}, f, indent=1)
    print("wrote", len(data), "arrays for", len(CASES), "cases")


if __name__ == "__main__":
    main()
