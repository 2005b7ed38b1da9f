#!/usr/bin/env python
"""Diagnostic: fused full-sweep kernels against the per-colour pass kernels, sweep by sweep."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from classicalspinmc.jl_b200 import _lib  # noqa: E402
from classicalspinmc.jl_b200._abi import FLAG_FUSED, FLAG_JIT, FLAG_NO_GRAPH, FLAG_NO_RESIDENT, ModelData  # noqa: E402
from classicalspinmc.jl_b200 import workloads as models  # noqa: E402

for name, uc, shape in (("square", models.square_heisenberg(), (64, 64)), ("honeycomb", models.kitaev_honeycomb(), (32, 32))):
    md = ModelData(uc, shape, 1.0)
    for label, orc, mc in (("or x2", 2, 0), ("metro x2", 0, 2)):
        res = []
        for extra in (FLAG_FUSED, 0):
            eng = _lib.Engine(md, seed=3, flags=FLAG_JIT | FLAG_NO_RESIDENT | FLAG_NO_GRAPH | extra)
            eng.randomize(5)
            eng.set_temperatures(1.0)
            l0 = eng.launches
            eng.cycles_async(1, orc, mc)
            eng.sync()
            res.append((eng.get_spins().copy(), eng.accepted()[0], eng.launches - l0))
        d = np.abs(res[0][0] - res[1][0])
        print(name, label, "launches", res[0][2], res[1][2], "acc", res[0][1], res[1][1], "max diff", d.max(),
              "n sites differing", int((d.max(axis=1) > 0).sum()), "of", d.shape[0])
