"""Lattice — host-side mirror of src/lattice.jl:4-345.

Public fields keep the reference's names (``S, spins, unit_cell, bc, size, shape, site_positions,
onsite, bilinear_sites, bilinear_matrices, cubic_sites, cubic_tensors, quartic_sites,
quartic_tensors, field``).  ``spins`` is a (3, N) Fortran-ordered float64 array — the same memory
as Julia's 3 x N ``Array{Float64,2}`` — so ``lat.spins[:, i]`` reads site i (0-based in Python).
The per-site tables are derived lazily in closed form (O(N * terms)); the reference's constructor
scans all N index tuples per slot (src/lattice.jl:196,228-229,273-275).

The device state (colour-major SoA spins, neighbour tables) lives behind ``lat._engine`` and is
created on first use; the host ``spins`` array stays the source of truth for a bare Lattice.
"""
from __future__ import annotations

import math

import numpy as np

from ._abi import ModelData
from .unit_cell import UnitCell


def random_spin_orientation(S, rng=None):
    """src/lattice.jl:306-311 — uniform point on the sphere of radius S."""
    rng = np.random.default_rng() if rng is None else rng
    phi = 2.0 * math.pi * rng.random()
    z = 2.0 * rng.random() - 1.0
    r = math.sqrt(1.0 - z * z)
    return (S * (r * math.cos(phi)), S * (r * math.sin(phi)), S * z)


def set_spin(spins, newspin, point):
    """src/lattice.jl:297-301 (``point`` is 1-based, as in the reference)."""
    spins[0, point - 1], spins[1, point - 1], spins[2, point - 1] = newspin


def get_spin(spins, point):
    """src/lattice.jl:293-295 (1-based)."""
    return (spins[0, point - 1], spins[1, point - 1], spins[2, point - 1])


def site_indices(shape, basis=1):
    """src/lattice.jl:29-33: sorted (basis, i1..iD) tuples, 1-based."""
    grids = np.indices((basis,) + tuple(shape)).reshape(len(shape) + 1, -1).T + 1
    return [tuple(int(v) for v in row) for row in grids]


def compute_site_positions(uc: UnitCell, size):
    """src/lattice.jl:38-51 -> (D, N) array."""
    D = uc.D
    nb = len(uc.basis)
    idx = np.indices((nb,) + tuple(size)).reshape(D + 1, -1)          # basis slowest, last dim fastest
    A = np.stack(uc.lattice_vectors, axis=1)                         # columns a_d
    pos = A @ idx[1:].astype(np.float64)
    pos += np.stack(uc.basis, axis=1)[:, idx[0]]
    return np.asfortranarray(pos)


class Lattice:
    def __init__(self, shape, uc: UnitCell, S=0.5, bc="periodic", initialCondition="random", rng=None):
        shape = tuple(int(s) for s in shape)
        if len(shape) != uc.D:
            raise ValueError("shape and unit cell dimension differ")
        if bc not in ("periodic", "open"):
            raise ValueError("Invalid boundary condition option")          # src/lattice.jl:107
        self._model = ModelData(uc, shape, S, bc)    # also adds the default basis site (:68-70)
        self.S = S
        self.unit_cell = uc
        self.bc = bc
        self.shape = shape
        self.size = self._model.n_sites
        N = self.size
        rng = np.random.default_rng() if rng is None else rng
        self.spins = np.zeros((3, N), order="F")
        if initialCondition in ("random", ":random"):                      # :76-79
            phi = 2.0 * math.pi * rng.random(N)
            z = 2.0 * rng.random(N) - 1.0
            r = np.sqrt(1.0 - z * z)
            self.spins[0], self.spins[1], self.spins[2] = S * r * np.cos(phi), S * r * np.sin(phi), S * z
        elif initialCondition in ("fm", ":fm"):                            # :80-85
            self.spins[:] = np.array(random_spin_orientation(S, rng))[:, None]
        else:
            raise ValueError("initialCondition must be 'random' or 'fm'")
        self.site_positions = compute_site_positions(uc, shape)
        self._tables = None
        self._engine = None

    # ---- lazily derived reference-layout tables ----------------------------------------------------
    def _basis_of_site(self):
        return np.repeat(np.arange(self._model.n_basis), self.size // self._model.n_basis)

    def _build_tables(self):
        if self._tables is not None:
            return self._tables
        from . import _lib
        md = self._model
        bil, cub, quar = _lib.reference_tables(md)
        b_of = self._basis_of_site() + 1
        t = {"bilinear_sites": bil, "cubic_sites": cub, "quartic_sites": quar,
             "field": md.field[b_of - 1].copy(), "onsite": md.onsite[b_of - 1].reshape(-1, 3, 3).copy()}
        # coupling "perspectives" (src/lattice.jl:184-194, 215-226, 251-271), null slots are zero
        mats = np.zeros((self.size, md.n2, 3, 3))
        for k, (b1, b2, M, off) in enumerate(self.unit_cell.bilinear):
            fwd = (b_of == b1) if b1 != b2 else (b_of == b1)
            bwd = (b_of == b2) & (b1 != b2)
            mats[fwd, k] = M
            mats[bwd, k] = M.T
            mats[bil[:, k] == 0, k] = 0.0
        t["bilinear_matrices"] = mats
        tens3 = np.zeros((self.size, md.n3, 3, 3, 3))
        for k, (b1, b2, b3, M, o2, o3) in enumerate(self.unit_cell.cubic):
            m1 = b_of == b1
            m2 = (b_of == b2) & ~m1
            m3 = (b_of == b3) & ~m1 & ~m2
            tens3[m1, k] = M
            tens3[m2, k] = np.transpose(M, (1, 0, 2))
            tens3[m3, k] = np.transpose(M, (2, 1, 0))
            tens3[cub[:, k, 0] == 0, k] = 0.0
        t["cubic_tensors"] = tens3
        tens4 = np.zeros((self.size, md.n4, 3, 3, 3, 3))
        for k, (b1, b2, b3, b4, M, o2, o3, o4) in enumerate(self.unit_cell.quartic):
            m1 = b_of == b1
            m2 = (b_of == b2) & ~m1
            m3 = (b_of == b3) & ~m1 & ~m2
            m4 = (b_of == b4) & ~m1 & ~m2 & ~m3
            tens4[m1, k] = M
            tens4[m2, k] = np.transpose(M, (1, 0, 2, 3))
            tens4[m3, k] = np.transpose(M, (2, 1, 0, 3))
            tens4[m4, k] = np.transpose(M, (3, 1, 2, 0))
            tens4[quar[:, k, 0] == 0, k] = 0.0
        t["quartic_tensors"] = tens4
        self._tables = t
        return t

    bilinear_sites = property(lambda self: self._build_tables()["bilinear_sites"])
    bilinear_matrices = property(lambda self: self._build_tables()["bilinear_matrices"])
    cubic_sites = property(lambda self: self._build_tables()["cubic_sites"])
    cubic_tensors = property(lambda self: self._build_tables()["cubic_tensors"])
    quartic_sites = property(lambda self: self._build_tables()["quartic_sites"])
    quartic_tensors = property(lambda self: self._build_tables()["quartic_tensors"])
    field = property(lambda self: self._build_tables()["field"])
    onsite = property(lambda self: self._build_tables()["onsite"])

    # ---- device mirror ----------------------------------------------------------------------------------
    def engine(self, **kw):
        """The single-replica device engine of this lattice (created on first use)."""
        if self._engine is None:
            from . import _lib
            self._engine = _lib.Engine(self._model, n_replicas=1, **kw)
        return self._engine

    def _spins_rows(self):
        """(N, 3) C-contiguous view of ``spins`` (zero-copy when spins is the usual (3, N) F array)."""
        s = self.spins
        if s.shape != (3, self.size):
            raise ValueError("lat.spins must have shape (3, N)")
        v = s.T
        return v if v.flags["C_CONTIGUOUS"] and v.dtype == np.float64 else np.ascontiguousarray(v, np.float64)

    def upload(self):
        self.engine().set_spins(self._spins_rows())

    def download(self):
        out = self.engine().get_spins()
        self.spins = np.asfortranarray(out.T)

    def copy(self):
        """deepcopy(lattice) as MonteCarlo() does (src/monte_carlo.jl:74): independent spins, shared model."""
        new = object.__new__(Lattice)
        new.__dict__.update(self.__dict__)
        new.spins = self.spins.copy(order="F")
        new._engine = None
        return new
