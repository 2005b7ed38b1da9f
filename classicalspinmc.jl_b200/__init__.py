"""B200-native sweep engine behind ClassicalSpinMC.jl's API (Python host mirror).

The compute path is libcsmc.so (hand-written sm_100a CUDA behind the C-ABI of include/csmc.h);
this package is the host side that plays the role of the reference's Julia layer.  Names follow
the reference's exports (src/ClassicalSpinMC.jl:6-35) with the trailing ``!`` dropped.
"""
from .unit_cell import (UnitCell, addBasisSite, addBilinear, addCubic, addOnSite, addQuartic,
                        addZeemanCoupling)
from .bravais import BreathingPyrochlore, FCC, Honeycomb, Pyrochlore, Square, Triangular

__all__ = [
    "UnitCell", "addBasisSite", "addBilinear", "addCubic", "addQuartic", "addZeemanCoupling",
    "addOnSite", "Triangular", "Square", "Honeycomb", "FCC", "Pyrochlore", "BreathingPyrochlore",
]
