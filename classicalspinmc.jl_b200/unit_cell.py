"""UnitCell and its builders — host-side mirror of the reference's src/unit_cell.jl:3-75.

Julia's ``addBilinear!`` etc. lose the ``!`` in Python (``addBilinear``); argument order, meaning
and the "all-zero couplings are silently dropped" rule (src/unit_cell.jl:38,49,60,72) are kept.
Basis indices stay 1-based exactly as in the reference, so user scripts translate line by line.
"""
from __future__ import annotations

import numpy as np


class UnitCell:
    """``UnitCell(a1, ..., aD)``; fields as src/unit_cell.jl:3-10."""

    def __init__(self, *lattice_vectors):
        if len(lattice_vectors) == 0:
            raise ValueError("UnitCell needs at least one lattice vector")
        self.lattice_vectors = tuple(np.asarray(a, dtype=np.float64).copy() for a in lattice_vectors)
        self.basis = []      # list of D-vectors
        self.field = []      # (b, h[3])
        self.onsite = []     # (b, M[3,3])
        self.bilinear = []   # (b1, b2, M[3,3], offset[D])
        self.cubic = []      # (b1, b2, b3, M[3,3,3], o2, o3)
        self.quartic = []    # (b1, b2, b3, b4, M[3,3,3,3], o2, o3, o4)

    @property
    def D(self) -> int:
        return len(self.lattice_vectors)

    def _offset(self, off):
        if off is None:
            return tuple([0] * self.D)
        off = tuple(int(o) for o in off)
        if len(off) != self.D:
            raise ValueError(f"offset {off} must have {self.D} components")
        return off


def _basis_index(b):
    if int(b) != b or int(b) < 1:
        raise ValueError("basis indices are 1-based integers (as in the reference)")
    return int(b)


def addBasisSite(uc: UnitCell, site):
    """src/unit_cell.jl:23-25"""
    uc.basis.append(np.asarray(site, dtype=np.float64).copy())


def addZeemanCoupling(uc: UnitCell, b1, h):
    """src/unit_cell.jl:30-32"""
    h = np.asarray(h, dtype=np.float64).reshape(3).copy()
    uc.field.append((_basis_index(b1), h))


def addOnSite(uc: UnitCell, b1, M):
    """src/unit_cell.jl:37-41"""
    M = np.asarray(M, dtype=np.float64).reshape(3, 3).copy()
    if np.any(M != 0):
        uc.onsite.append((_basis_index(b1), M))


def addBilinear(uc: UnitCell, b1, b2, M, offset=None):
    """src/unit_cell.jl:46-52"""
    M = np.asarray(M, dtype=np.float64).reshape(3, 3).copy()
    if np.any(M != 0):
        uc.bilinear.append((_basis_index(b1), _basis_index(b2), M, uc._offset(offset)))


def addCubic(uc: UnitCell, b1, b2, b3, M, o2=None, o3=None):
    """src/unit_cell.jl:57-63"""
    M = np.asarray(M, dtype=np.float64).reshape(3, 3, 3).copy()
    if np.any(M != 0):
        uc.cubic.append((_basis_index(b1), _basis_index(b2), _basis_index(b3), M,
                         uc._offset(o2), uc._offset(o3)))


def addQuartic(uc: UnitCell, b1, b2, b3, b4, M, o2=None, o3=None, o4=None):
    """src/unit_cell.jl:68-75"""
    M = np.asarray(M, dtype=np.float64).reshape(3, 3, 3, 3).copy()
    if np.any(M != 0):
        uc.quartic.append((_basis_index(b1), _basis_index(b2), _basis_index(b3), _basis_index(b4), M,
                           uc._offset(o2), uc._offset(o3), uc._offset(o4)))
