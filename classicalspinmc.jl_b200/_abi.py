"""ctypes view of include/csmc.h (struct layouts + model marshalling).

This is the Python stand-in for the Julia ``ccall`` layer (julia/ClassicalSpinMC/src/libcsmc.jl):
the same structs, filled from the same ``UnitCell`` / ``Lattice`` data.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

MAX_DIM = 3


class CsmcModel(C.Structure):
    _fields_ = [
        ("dim", C.c_int32),
        ("shape", C.c_int32 * MAX_DIM),
        ("n_basis", C.c_int32),
        ("periodic", C.c_int32),
        ("S", C.c_double),
        ("field", C.POINTER(C.c_double)),
        ("onsite", C.POINTER(C.c_double)),
        ("n_bilinear", C.c_int32),
        ("bil_basis", C.POINTER(C.c_int32)),
        ("bil_offset", C.POINTER(C.c_int32)),
        ("bil_matrix", C.POINTER(C.c_double)),
        ("n_cubic", C.c_int32),
        ("cub_basis", C.POINTER(C.c_int32)),
        ("cub_offset", C.POINTER(C.c_int32)),
        ("cub_tensor", C.POINTER(C.c_double)),
        ("n_quartic", C.c_int32),
        ("quar_basis", C.POINTER(C.c_int32)),
        ("quar_offset", C.POINTER(C.c_int32)),
        ("quar_tensor", C.POINTER(C.c_double)),
    ]


class CsmcOpts(C.Structure):
    _fields_ = [
        ("device", C.c_int32),
        ("n_replicas", C.c_int32),
        ("seed", C.c_uint64),
        ("stream", C.c_void_p),
        ("replica_base", C.c_int32),
        ("flags", C.c_int32),
    ]


class CsmcPtParams(C.Structure):
    _fields_ = [
        ("t_thermalization", C.c_int64),
        ("t_measurement", C.c_int64),
        ("probe_rate", C.c_int32),
        ("swap_rate", C.c_int32),
        ("overrelaxation_rate", C.c_int32),
        ("algorithm", C.c_int32),
    ]


FLAG_FORCE_GENERIC = 1
FLAG_NO_GRAPH = 2
FLAG_JIT = 4
FLAG_NO_JIT = 8
FLAG_PDL = 16
FLAG_NO_RESIDENT = 32
FLAG_NO_AUTOTUNE = 64
FLAG_FUSED = 128
FLAG_SKEW = 256
FLAG_NO_PERSIST = 512
FLAG_PERSIST = 1024


def resolve_field_onsite(uc):
    """Per-basis Zeeman vector and on-site matrix exactly as src/lattice.jl:117-140 resolves them.

    The reference indexes the *term list* by basis number (``f_indices[i]``/``f_[i]``), which only
    works when terms are given for a leading subset {1..k} of the basis sites (in any order);
    other usages raise a BoundsError or read uninitialised memory in Julia.  Those cases raise
    here instead of inventing values.
    """
    nb = len(uc.basis)
    field = np.zeros((nb, 3))
    onsite = np.zeros((nb, 9))
    f_idx = [t[0] for t in uc.field]
    o_idx = [t[0] for t in uc.onsite]
    f_set = [False] * nb
    o_set = [False] * nb
    for i in range(1, nb + 1):
        if i not in f_idx:
            f_set[i - 1] = True                                   # :129-130
        else:
            if i > len(f_idx):
                raise IndexError("Zeeman terms must cover basis sites 1..k (src/lattice.jl:132 "
                                 "indexes the term list by basis number)")
            b = f_idx[i - 1]
            if not 1 <= b <= nb:
                raise IndexError(f"Zeeman basis index {b} out of range")
            field[b - 1] = uc.field[i - 1][1]                     # :132
            f_set[b - 1] = True
        if i not in o_idx:
            o_set[i - 1] = True                                   # :135-136
        else:
            if i > len(o_idx):
                raise IndexError("on-site terms must cover basis sites 1..k (src/lattice.jl:138)")
            b = o_idx[i - 1]
            if not 1 <= b <= nb:
                raise IndexError(f"on-site basis index {b} out of range")
            onsite[b - 1] = np.asarray(uc.onsite[i - 1][1]).reshape(9)   # :138
            o_set[b - 1] = True
    if not all(f_set) or not all(o_set):
        raise ValueError("field/onsite assignment leaves a basis site undefined in the reference "
                         "(src/lattice.jl:122,126 allocate `undef`); give terms for sites 1..k")
    return field, onsite


class ModelData:
    """Owns the numpy buffers a CsmcModel points into."""

    def __init__(self, uc, shape, S, bc="periodic"):
        D = uc.D
        if D > MAX_DIM:
            raise ValueError(f"at most {MAX_DIM} lattice dimensions are supported")
        shape = tuple(int(s) for s in shape)
        if len(shape) != D or any(s < 1 for s in shape):
            raise ValueError("shape must have one positive entry per lattice vector")
        if bc not in ("periodic", "open"):
            raise ValueError("Invalid boundary condition option")   # src/lattice.jl:107
        if len(uc.basis) == 0:                                      # src/lattice.jl:68-70
            uc.basis.append(np.zeros(D))
        nb = len(uc.basis)
        self.D, self.shape, self.S, self.bc, self.n_basis = D, shape, float(S), bc, nb
        self.field, self.onsite = resolve_field_onsite(uc)

        def chk(b):
            if not 1 <= b <= nb:
                raise IndexError(f"basis index {b} out of range 1..{nb}")
            return b

        n2, n3, n4 = len(uc.bilinear), len(uc.cubic), len(uc.quartic)
        self.bil_basis = np.zeros((max(n2, 1), 2), np.int32)
        self.bil_offset = np.zeros((max(n2, 1), D), np.int32)
        self.bil_matrix = np.zeros((max(n2, 1), 9))
        for t, (b1, b2, M, off) in enumerate(uc.bilinear):
            self.bil_basis[t] = (chk(b1), chk(b2))
            self.bil_offset[t] = off
            self.bil_matrix[t] = np.asarray(M).reshape(9)          # row-major m11 m12 ...
        self.cub_basis = np.zeros((max(n3, 1), 3), np.int32)
        self.cub_offset = np.zeros((max(n3, 1), 2, D), np.int32)
        self.cub_tensor = np.zeros((max(n3, 1), 27))
        for t, (b1, b2, b3, M, o2, o3) in enumerate(uc.cubic):
            self.cub_basis[t] = (chk(b1), chk(b2), chk(b3))
            self.cub_offset[t, 0], self.cub_offset[t, 1] = o2, o3
            self.cub_tensor[t] = np.asarray(M).ravel(order="F")    # Julia column-major
        self.quar_basis = np.zeros((max(n4, 1), 4), np.int32)
        self.quar_offset = np.zeros((max(n4, 1), 3, D), np.int32)
        self.quar_tensor = np.zeros((max(n4, 1), 81))
        for t, (b1, b2, b3, b4, M, o2, o3, o4) in enumerate(uc.quartic):
            self.quar_basis[t] = (chk(b1), chk(b2), chk(b3), chk(b4))
            self.quar_offset[t, 0], self.quar_offset[t, 1], self.quar_offset[t, 2] = o2, o3, o4
            self.quar_tensor[t] = np.asarray(M).ravel(order="F")
        self.n2, self.n3, self.n4 = n2, n3, n4
        self.n_sites = int(np.prod(shape)) * nb

        dp = lambda a: np.ascontiguousarray(a).ctypes.data_as(C.POINTER(C.c_double))
        ip = lambda a: np.ascontiguousarray(a).ctypes.data_as(C.POINTER(C.c_int32))
        m = CsmcModel()
        m.dim = D
        for d in range(MAX_DIM):
            m.shape[d] = shape[d] if d < D else 1
        m.n_basis = nb
        m.periodic = 1 if bc == "periodic" else 0
        m.S = float(S)
        m.field, m.onsite = dp(self.field), dp(self.onsite)
        m.n_bilinear, m.bil_basis, m.bil_offset, m.bil_matrix = n2, ip(self.bil_basis), ip(self.bil_offset), dp(self.bil_matrix)
        m.n_cubic, m.cub_basis, m.cub_offset, m.cub_tensor = n3, ip(self.cub_basis), ip(self.cub_offset), dp(self.cub_tensor)
        m.n_quartic, m.quar_basis, m.quar_offset, m.quar_tensor = n4, ip(self.quar_basis), ip(self.quar_offset), dp(self.quar_tensor)
        self.struct = m
