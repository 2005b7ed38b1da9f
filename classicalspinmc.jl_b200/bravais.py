"""Bravais presets — mirror of src/bravais.jl:1-51."""
import math

import numpy as np

from .unit_cell import UnitCell, addBasisSite


def Triangular() -> UnitCell:
    """src/bravais.jl:1-8"""
    a1 = np.array([1.0, 0.0])
    a2 = math.cos(math.pi / 3) * np.array([1.0, 0.0]) + math.sin(math.pi / 3) * np.array([0.0, 1.0])
    return UnitCell(a1, a2)


def Square() -> UnitCell:
    """src/bravais.jl:10-16"""
    return UnitCell(np.array([1.0, 0.0]), np.array([0.0, 1.0]))


def FCC() -> UnitCell:
    """src/bravais.jl:18-25"""
    return UnitCell(0.5 * np.array([0.0, 1, 1]), 0.5 * np.array([1.0, 0, 1]), 0.5 * np.array([1.0, 1, 0]))


def Pyrochlore() -> UnitCell:
    """src/bravais.jl:27-34"""
    uc = FCC()
    for site in ([1, 1, 1], [1, -1, -1], [-1, 1, -1], [-1, -1, 1]):
        addBasisSite(uc, np.array(site, dtype=np.float64) / 8)
    return uc


def BreathingPyrochlore(a: float = 1.01) -> UnitCell:
    """src/bravais.jl:36-43"""
    uc = FCC()
    for site in ([1, 1, 1], [1, -1, -1], [-1, 1, -1], [-1, -1, 1]):
        addBasisSite(uc, a * np.array(site, dtype=np.float64) / 8)
    return uc


def Honeycomb() -> UnitCell:
    """src/bravais.jl:45-51"""
    uc = Triangular()
    addBasisSite(uc, np.array([0.0, 0.0]))
    addBasisSite(uc, np.array([0.0, 1.0]) / math.sqrt(3))
    return uc
