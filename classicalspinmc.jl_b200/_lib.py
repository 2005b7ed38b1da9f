"""Loader and ctypes prototypes for libcsmc.so (include/csmc.h) plus a thin ``Engine`` wrapper.

There is no CPU fallback: if the shared library is missing or a call fails, a ``CsmcError`` is
raised.  ``build()`` compiles the library in-tree with nvcc for sm_100a (works without a GPU).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from ._abi import CsmcModel, CsmcOpts, CsmcPtParams, ModelData

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libcsmc.so")
_LIB = None

# every symbol include/csmc.h declares (tests check the .so exports all of them)
EXPORTS = [
    "csmc_version", "csmc_last_error", "csmc_create", "csmc_destroy", "csmc_plan", "csmc_reference_tables",
    "csmc_n_sites",
    "csmc_n_replicas", "csmc_n_colours", "csmc_get_colouring", "csmc_is_structured",
    "csmc_kernel_mode", "csmc_autotune_report", "csmc_sweep_groups", "csmc_jit_check", "csmc_launch_count", "csmc_get_tables", "csmc_set_spins", "csmc_get_spins",
    "csmc_randomize_spins", "csmc_local_field", "csmc_local_field_all", "csmc_site_energy_all",
    "csmc_total_energy", "csmc_magnetization", "csmc_structure_factor", "csmc_overrelax", "csmc_deterministic",
    "csmc_metropolis", "csmc_metropolis_cone", "csmc_anneal_temperature", "csmc_anneal_temperature_cone", "csmc_set_temperatures",
    "csmc_set_sigma", "csmc_get_sigma",
    "csmc_cycles_async", "csmc_sync", "csmc_get_accepted", "csmc_pt_init", "csmc_comm_unique_id",
    "csmc_comm_init", "csmc_comm_mode", "csmc_replica_blocks", "csmc_persist_info", "csmc_persist_check", "csmc_kernel_costs", "csmc_skew_schedule", "csmc_skew_info", "csmc_skew_geometry", "csmc_pt_run", "csmc_pt_exchange", "csmc_pt_get_slots", "csmc_pt_get_series",
    "csmc_pt_get_stats", "csmc_pt_set_momenta", "csmc_pt_get_ssf",
]


class CsmcError(RuntimeError):
    pass


def build(force: bool = False) -> str:
    """Compile csrc/ -> libcsmc.so (nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo)."""
    src_dir = os.path.join(_HERE, "csrc")
    srcs = [os.path.join(src_dir, f) for f in os.listdir(src_dir) if f.endswith((".cu", ".cuh", ".cpp", ".h"))]
    srcs.append(os.path.join(os.path.dirname(_HERE), "include", "csmc.h"))
    stale = (not os.path.exists(_SO)) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs)
    if force or stale:
        subprocess.check_call(["make", "-C", src_dir, "-s"] + (["-B"] if force else []))
    return _SO


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(_SO):
        raise CsmcError(f"{_SO} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(there is no CPU fallback)")
    L = C.CDLL(_SO)
    vp, i32, i64, u64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.c_double
    P = C.POINTER
    L.csmc_version.restype = i32
    L.csmc_last_error.restype = C.c_char_p
    L.csmc_last_error.argtypes = [vp]
    L.csmc_create.argtypes = [P(CsmcModel), P(CsmcOpts), P(vp)]
    L.csmc_destroy.argtypes = [vp]
    L.csmc_plan.argtypes = [P(CsmcModel), i32, vp, P(i32), P(i32), vp]
    L.csmc_reference_tables.argtypes = [P(CsmcModel), vp, vp, vp]
    L.csmc_n_sites.argtypes = [vp, P(i64)]
    L.csmc_n_replicas.argtypes = [vp, P(i32)]
    L.csmc_n_colours.argtypes = [vp, P(i32)]
    L.csmc_get_colouring.argtypes = [vp, vp]
    L.csmc_is_structured.argtypes = [vp, P(i32)]
    L.csmc_launch_count.argtypes = [vp, P(i64)]
    L.csmc_kernel_mode.argtypes = [vp, P(i32)]
    L.csmc_autotune_report.argtypes = [vp, vp, P(i32)]
    L.csmc_sweep_groups.argtypes = [vp, P(i32), vp]
    L.csmc_replica_blocks.argtypes = [vp, P(i32), vp]
    L.csmc_persist_info.argtypes = [vp, P(i32), vp, P(i32), P(i32), vp]
    L.csmc_kernel_costs.argtypes = [vp, P(dbl), P(dbl)]
    L.csmc_persist_check.argtypes = [P(CsmcModel), i32, i32, i32, i32, vp, i64, P(i64), vp, i64, vp]
    L.csmc_skew_schedule.argtypes = [i32, i32, i32, i32, vp, i64, P(i64)]
    L.csmc_skew_info.argtypes = [vp, P(i32), P(i32), P(i32), P(i32)]
    L.csmc_skew_geometry.argtypes = [P(CsmcModel), P(i32), P(i32), P(i32), P(i32)]
    L.csmc_jit_check.argtypes = [P(CsmcModel), i32, vp, i64, P(i64), vp, i64]
    L.csmc_get_tables.argtypes = [vp, vp, vp, vp]
    L.csmc_set_spins.argtypes = [vp, i32, vp]
    L.csmc_get_spins.argtypes = [vp, i32, vp]
    L.csmc_randomize_spins.argtypes = [vp, u64]
    L.csmc_local_field.argtypes = [vp, i32, i64, vp]
    L.csmc_local_field_all.argtypes = [vp, i32, vp]
    L.csmc_site_energy_all.argtypes = [vp, i32, vp]
    L.csmc_total_energy.argtypes = [vp, vp]
    L.csmc_magnetization.argtypes = [vp, vp]
    L.csmc_structure_factor.argtypes = [vp, i32, vp, vp, vp, i64, vp]
    L.csmc_overrelax.argtypes = [vp, i32]
    L.csmc_deterministic.argtypes = [vp, i32]
    L.csmc_metropolis.argtypes = [vp, vp, i32, vp]
    L.csmc_metropolis_cone.argtypes = [vp, vp, vp, i32, i32, vp]
    L.csmc_anneal_temperature.argtypes = [vp, vp, i64, i32, vp]
    L.csmc_anneal_temperature_cone.argtypes = [vp, vp, vp, i32, i64, i32, vp]
    L.csmc_set_temperatures.argtypes = [vp, vp]
    L.csmc_set_sigma.argtypes = [vp, vp]
    L.csmc_get_sigma.argtypes = [vp, vp]
    L.csmc_cycles_async.argtypes = [vp, i64, i32, i32]
    L.csmc_sync.argtypes = [vp]
    L.csmc_get_accepted.argtypes = [vp, vp, i32]
    L.csmc_pt_init.argtypes = [vp, i32, vp]
    L.csmc_comm_unique_id.argtypes = [vp]
    L.csmc_comm_init.argtypes = [vp, i32, i32, vp]
    L.csmc_comm_mode.argtypes = [vp, P(i32)]
    L.csmc_pt_run.argtypes = [vp, P(CsmcPtParams), i64, i64]
    L.csmc_pt_exchange.argtypes = [vp, i32, vp]
    L.csmc_pt_get_slots.argtypes = [vp, vp]
    L.csmc_pt_get_series.argtypes = [vp, P(i64), vp, vp]
    L.csmc_pt_get_stats.argtypes = [vp, vp, vp]
    L.csmc_pt_set_momenta.argtypes = [vp, vp, vp, vp, i64]
    L.csmc_pt_get_ssf.argtypes = [vp, vp, P(i64)]
    for name in EXPORTS:
        if name not in ("csmc_last_error",):
            getattr(L, name).restype = i32
    _LIB = L
    return L


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f64(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None and a.shape != tuple(shape):
        raise ValueError(f"expected shape {shape}, got {a.shape}")
    return a


def comm_unique_id() -> bytes:
    buf = (C.c_uint8 * 128)()
    rc = lib().csmc_comm_unique_id(buf)
    if rc:
        raise CsmcError(f"csmc_comm_unique_id failed ({rc}): {lib().csmc_last_error(None).decode()}")
    return bytes(buf)


def plan(model: ModelData, flags: int = 0):
    """Host-only colouring / layout plan (csmc_plan): (colour[N], n_colours, structured, storage_pos[N])."""
    L = lib()
    col = np.zeros(model.n_sites, np.int32)
    pos = np.zeros(model.n_sites, np.int32)
    nc, st = C.c_int32(), C.c_int32()
    rc = L.csmc_plan(C.byref(model.struct), flags, _p(col), C.byref(nc), C.byref(st), _p(pos))
    if rc:
        raise CsmcError(f"csmc_plan failed ({rc}): {L.csmc_last_error(None).decode()}")
    return col, nc.value, bool(st.value), pos


def jit_check(model: ModelData, compile: bool = True):
    """Host-only: generate (and optionally NVRTC-compile for sm_100a) the specialised kernel source of
    ``model``.  Returns (source, log); raises CsmcError when generation or compilation fails."""
    L = lib()
    n = C.c_int64(0)
    cap = 1 << 22
    src = C.create_string_buffer(cap)
    log = C.create_string_buffer(1 << 16)
    rc = L.csmc_jit_check(C.byref(model.struct), int(compile), src, cap, C.byref(n), log, 1 << 16)
    if rc:
        raise CsmcError(f"csmc_jit_check failed ({rc}): {L.csmc_last_error(None).decode()}")
    return src.value.decode(), log.value.decode()


def persist_check(model: ModelData, n_replicas: int = 1, n_sms: int = 0, smem_max: int = 0, compile: bool = True):
    """Host-only: plan (and optionally NVRTC-compile for sm_100a) the tile-resident persistent kernel of ``model``.
    Returns (info dict, source, log); info["usable"] is False when no tiling fits."""
    L = lib()
    n = C.c_int64(0)
    cap = 1 << 23
    src = C.create_string_buffer(cap)
    log = C.create_string_buffer(1 << 16)
    info = (C.c_int32 * 8)()
    rc = L.csmc_persist_check(C.byref(model.struct), n_replicas, n_sms, smem_max, int(compile), src, cap, C.byref(n), log, 1 << 16, info)
    if rc:
        raise CsmcError(f"csmc_persist_check failed ({rc}): {L.csmc_last_error(None).decode()}")
    keys = ("usable", "tiles", "g0", "g1", "w0", "w1", "replicas_per_launch", "smem")
    d = dict(zip(keys, [int(v) for v in info]))
    d["usable"] = bool(d["usable"])
    return d, src.value.decode(), log.value.decode()


def skew_schedule(n_rows: int, n_passes: int, reach: int, budget_rows: int):
    """Host-only launch plan of the time-skewed strips (csmc_skew_schedule): int32 array (n, 3) of
    (pass, first tile row, tile rows); empty when not applicable."""
    L = lib()
    n = C.c_int64(0)
    rc = L.csmc_skew_schedule(n_rows, n_passes, reach, budget_rows, None, 0, C.byref(n))
    if rc:
        raise CsmcError(f"csmc_skew_schedule failed ({rc}): {L.csmc_last_error(None).decode()}")
    out = np.zeros((n.value, 3), np.int32)
    if n.value:
        L.csmc_skew_schedule(n_rows, n_passes, reach, budget_rows, _p(out), n.value, C.byref(n))
    return out


def skew_geometry(model: ModelData):
    """Host-only: (usable, CTA-tile rows along dimension 0, reach in rows, CTA tiles per row) of the time-skewed strips."""
    v = [C.c_int32() for _ in range(4)]
    rc = lib().csmc_skew_geometry(C.byref(model.struct), *[C.byref(x) for x in v])
    if rc:
        raise CsmcError(f"csmc_skew_geometry failed ({rc}): {lib().csmc_last_error(None).decode()}")
    return bool(v[0].value), v[1].value, v[2].value, v[3].value


def reference_tables(model: ModelData):
    """Host-only closed-form neighbour tables in the reference's layout (1-based, 0 == null)."""
    L = lib()
    N = model.n_sites
    bil = np.zeros((N, model.n2), np.int64)
    cub = np.zeros((N, model.n3, 2), np.int64)
    quar = np.zeros((N, model.n4, 3), np.int64)
    rc = L.csmc_reference_tables(C.byref(model.struct), _p(bil), _p(cub), _p(quar))
    if rc:
        raise CsmcError(f"csmc_reference_tables failed ({rc}): {L.csmc_last_error(None).decode()}")
    return bil, cub, quar


class Engine:
    """Owns one ``csmc_handle`` (device state of ``n_replicas`` replicas of one lattice model).

    Spin arrays cross this boundary as (N, 3) C-contiguous float64, the memory layout of the
    reference's ``lattice.spins`` (Julia 3 x N column-major)."""

    def __init__(self, model: ModelData, n_replicas: int = 1, seed: int = 12345, device: int = 0,
                 stream: int | None = None, replica_base: int = 0, flags: int = 0):
        self._L = lib()
        self.model = model
        opts = CsmcOpts(device=device, n_replicas=n_replicas, seed=seed, stream=stream,
                        replica_base=replica_base, flags=flags)
        h = C.c_void_p()
        rc = self._L.csmc_create(C.byref(model.struct), C.byref(opts), C.byref(h))
        if rc:
            raise CsmcError(f"csmc_create failed ({rc}): {self._L.csmc_last_error(None).decode()}")
        self._h = h
        self.n_replicas = n_replicas
        self.replica_base = replica_base
        self.seed = seed
        n = C.c_int64()
        self._ck(self._L.csmc_n_sites(self._h, C.byref(n)))
        self.N = n.value

    # -- plumbing -----------------------------------------------------------------------------
    def _ck(self, rc):
        if rc:
            raise CsmcError(f"libcsmc error {rc}: {self._L.csmc_last_error(self._h).decode()}")

    def close(self):
        if getattr(self, "_h", None):
            self._L.csmc_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- introspection ------------------------------------------------------------------------
    @property
    def n_colours(self):
        c = C.c_int32()
        self._ck(self._L.csmc_n_colours(self._h, C.byref(c)))
        return c.value

    @property
    def structured(self):
        c = C.c_int32()
        self._ck(self._L.csmc_is_structured(self._h, C.byref(c)))
        return bool(c.value)

    @property
    def kernel_mode(self):
        """0 explicit-table, 1 arithmetic-neighbour (ahead of time), 2 runtime-specialised (NVRTC)."""
        c = C.c_int32()
        self._ck(self._L.csmc_kernel_mode(self._h, C.byref(c)))
        return c.value

    def autotune_report(self):
        """(ms without PDL, ms with PDL, selected) of the launch-mode autotune at create (zeros if skipped)."""
        ms = (C.c_float * 2)()
        sel = C.c_int32()
        self._ck(self._L.csmc_autotune_report(self._h, ms, C.byref(sel)))
        return float(ms[0]), float(ms[1]), bool(sel.value)

    def sweep_groups(self):
        """(replica groups in use, (ms with 1, 2, 4 groups) of the create-time probe; zeros if not measured)."""
        ms = (C.c_float * 3)()
        g = C.c_int32()
        self._ck(self._L.csmc_sweep_groups(self._h, C.byref(g), ms))
        return int(g.value), tuple(float(v) for v in ms)

    def skew_info(self):
        """(usable, tile rows, reach in tile rows, L2 budget in tile rows) of the time-skewed strips (CSMC_FLAG_SKEW)."""
        v = [C.c_int32() for _ in range(4)]
        self._ck(self._L.csmc_skew_info(self._h, *[C.byref(x) for x in v]))
        return bool(v[0].value), v[1].value, v[2].value, v[3].value

    def persist_info(self):
        """(CTA tiles per replica of the tile-resident kernel (0: not in use), (tiles along dim 0, dim 1), replicas per
        launch, shared memory per CTA, (ms pass kernels, ms persistent) of the create-time probe)."""
        t, r, sm = C.c_int32(), C.c_int32(), C.c_int32()
        g = (C.c_int32 * 2)()
        ms = (C.c_float * 2)()
        self._ck(self._L.csmc_persist_info(self._h, C.byref(t), g, C.byref(r), C.byref(sm), ms))
        return int(t.value), (int(g[0]), int(g[1])), int(r.value), int(sm.value), (float(ms[0]), float(ms[1]))

    def kernel_costs(self):
        """(fp64 flops per overrelaxation site update of the specialised kernels, algorithmic bytes per update)."""
        f, b = C.c_double(), C.c_double()
        self._ck(self._L.csmc_kernel_costs(self._h, C.byref(f), C.byref(b)))
        return float(f.value), float(b.value)

    def replica_blocks(self):
        """(replica blocks in use, (ms unblocked, ms blocked) of the create-time probe; zeros if not measured)."""
        ms = (C.c_float * 2)()
        b = C.c_int32()
        self._ck(self._L.csmc_replica_blocks(self._h, C.byref(b), ms))
        return int(b.value), tuple(float(v) for v in ms)

    @property
    def launches(self):
        c = C.c_int64()
        self._ck(self._L.csmc_launch_count(self._h, C.byref(c)))
        return c.value

    def colouring(self):
        col = np.zeros(self.N, np.int32)
        self._ck(self._L.csmc_get_colouring(self._h, _p(col)))
        return col

    def colour_order(self):
        """1-based site visiting order of one colour-ordered sweep."""
        return (np.argsort(self.colouring(), kind="stable") + 1).astype(np.int64)

    def tables(self):
        md = self.model
        bil = np.zeros((self.N, md.n2), np.int64)
        cub = np.zeros((self.N, md.n3, 2), np.int64)
        quar = np.zeros((self.N, md.n4, 3), np.int64)
        self._ck(self._L.csmc_get_tables(self._h, _p(bil), _p(cub), _p(quar)))
        return bil, cub, quar

    # -- state --------------------------------------------------------------------------------
    def set_spins(self, spins, replica=0):
        s = _f64(spins, (self.N, 3))
        self._ck(self._L.csmc_set_spins(self._h, replica, _p(s)))

    def get_spins(self, replica=0, out=None):
        s = np.empty((self.N, 3)) if out is None else out
        self._ck(self._L.csmc_get_spins(self._h, replica, _p(s)))
        return s

    def randomize(self, seed):
        self._ck(self._L.csmc_randomize_spins(self._h, seed))

    # -- Hamiltonian --------------------------------------------------------------------------
    def local_field(self, site, replica=0):
        out = np.zeros(3)
        self._ck(self._L.csmc_local_field(self._h, replica, site, _p(out)))
        return out

    def local_field_all(self, replica=0):
        out = np.zeros((self.N, 3))
        self._ck(self._L.csmc_local_field_all(self._h, replica, _p(out)))
        return out

    def site_energy_all(self, replica=0):
        out = np.zeros(self.N)
        self._ck(self._L.csmc_site_energy_all(self._h, replica, _p(out)))
        return out

    def total_energy(self):
        E = np.zeros(self.n_replicas)
        self._ck(self._L.csmc_total_energy(self._h, _p(E)))
        return E

    def magnetization_vector(self):
        M = np.zeros((self.n_replicas, 3))
        self._ck(self._L.csmc_magnetization(self._h, _p(M)))
        return M

    def structure_factor(self, lattice_vectors, basis, ks, replica=0):
        """lattice_vectors: sequence of D vectors a_d; basis: sequence of n_basis D-vectors; ks: (D, N_k).
        Returns Suv (9, N_k) as compute_equal_time_correlations does."""
        A = np.ascontiguousarray(np.stack([np.asarray(a, dtype=np.float64) for a in lattice_vectors]))   # row d = a_d == column-major D x D
        B = np.ascontiguousarray(np.stack([np.asarray(b, dtype=np.float64) for b in basis]))
        K = np.ascontiguousarray(np.asarray(ks, dtype=np.float64).T)                                      # (N_k, D) == column-major D x N_k
        out = np.zeros((K.shape[0], 9))
        self._ck(self._L.csmc_structure_factor(self._h, replica, _p(A), _p(B), _p(K), K.shape[0], _p(out)))
        return np.ascontiguousarray(out.T)

    # -- sweeps -------------------------------------------------------------------------------
    def _T(self, T):
        T = np.broadcast_to(np.asarray(T, dtype=np.float64), (self.n_replicas,))
        return np.ascontiguousarray(T)

    def overrelax(self, n_sweeps=1):
        self._ck(self._L.csmc_overrelax(self._h, n_sweeps))

    def deterministic(self, n_sweeps=1):
        self._ck(self._L.csmc_deterministic(self._h, n_sweeps))

    def metropolis(self, T, n_sweeps=1):
        acc = np.zeros(self.n_replicas)
        self._ck(self._L.csmc_metropolis(self._h, _p(self._T(T)), n_sweeps, _p(acc)))
        return acc

    def metropolis_cone(self, T, sigma, adapt=False, n_sweeps=1):
        acc = np.zeros(self.n_replicas)
        sig = self._T(sigma).copy()
        self._ck(self._L.csmc_metropolis_cone(self._h, _p(self._T(T)), _p(sig), int(adapt), n_sweeps, _p(acc)))
        return acc, sig

    def anneal_temperature(self, T, t_thermalization, overrelaxation_rate):
        acc = np.zeros(self.n_replicas)
        self._ck(self._L.csmc_anneal_temperature(self._h, _p(self._T(T)), t_thermalization,
                                                 overrelaxation_rate, _p(acc)))
        return acc

    def anneal_temperature_cone(self, T, sigma, adapt, t_thermalization, overrelaxation_rate):
        acc = np.zeros(self.n_replicas)
        sig = self._T(sigma).copy()
        self._ck(self._L.csmc_anneal_temperature_cone(self._h, _p(self._T(T)), _p(sig), int(adapt), t_thermalization,
                                                      overrelaxation_rate, _p(acc)))
        return acc, sig

    def set_temperatures(self, T):
        self._ck(self._L.csmc_set_temperatures(self._h, _p(self._T(T))))

    def set_sigma(self, sigma):
        self._ck(self._L.csmc_set_sigma(self._h, _p(self._T(sigma))))

    def get_sigma(self):
        s = np.zeros(self.n_replicas)
        self._ck(self._L.csmc_get_sigma(self._h, _p(s)))
        return s

    def cycles_async(self, n_cycles, or_per_cycle, metro_per_cycle):
        self._ck(self._L.csmc_cycles_async(self._h, n_cycles, or_per_cycle, metro_per_cycle))

    def sync(self):
        self._ck(self._L.csmc_sync(self._h))

    def accepted(self, reset=False):
        acc = np.zeros(self.n_replicas)
        self._ck(self._L.csmc_get_accepted(self._h, _p(acc), int(reset)))
        return acc

    # -- parallel tempering -------------------------------------------------------------------
    def comm_init(self, n_ranks, rank, unique_id: bytes):
        buf = (C.c_uint8 * 128).from_buffer_copy(unique_id)
        self._ck(self._L.csmc_comm_init(self._h, n_ranks, rank, buf))

    def comm_mode(self) -> int:
        """0 no communicator, 1 NCCL collectives, 2 / 3 stores into peer memory (CSMC_PEER_GATHER=1 / 2)."""
        m = C.c_int32()
        self._ck(self._L.csmc_comm_mode(self._h, C.byref(m)))
        return m.value

    def pt_init(self, T_all):
        T_all = _f64(T_all)
        self.n_slots = len(T_all)
        self._ck(self._L.csmc_pt_init(self._h, len(T_all), _p(T_all)))

    def pt_run(self, params: dict, sweep_begin, sweep_end):
        p = CsmcPtParams(params["t_thermalization"], params["t_measurement"], params["probe_rate"],
                         params["swap_rate"], params["overrelaxation_rate"], int(params.get("algorithm", 0)))
        self._ck(self._L.csmc_pt_run(self._h, C.byref(p), sweep_begin, sweep_end))

    def pt_exchange(self, parity):
        acc = np.zeros(self.n_slots, np.int32)
        self._ck(self._L.csmc_pt_exchange(self._h, parity, _p(acc)))
        return acc

    def pt_slots(self):
        s = np.zeros(self.n_slots, np.int32)
        self._ck(self._L.csmc_pt_get_slots(self._h, _p(s)))
        return s

    def pt_series(self):
        n = C.c_int64(0)
        self._ck(self._L.csmc_pt_get_series(self._h, C.byref(n), None, None))
        cnt = n.value
        E = np.zeros((cnt, self.n_slots)); M = np.zeros((cnt, self.n_slots))
        n = C.c_int64(cnt)
        self._ck(self._L.csmc_pt_get_series(self._h, C.byref(n), _p(E), _p(M)))
        return E, M

    def pt_set_momenta(self, lattice_vectors, basis, ks):
        A = np.ascontiguousarray(np.stack([np.asarray(a, dtype=np.float64) for a in lattice_vectors]))
        B = np.ascontiguousarray(np.stack([np.asarray(b, dtype=np.float64) for b in basis]))
        K = np.ascontiguousarray(np.asarray(ks, dtype=np.float64).T)
        self._n_k = K.shape[0]
        self._ck(self._L.csmc_pt_set_momenta(self._h, _p(A), _p(B), _p(K), K.shape[0]))

    def pt_ssf(self):
        """(sums[n_slots, 9, n_k], n_probes): this rank's per-slot structure-factor sums."""
        sums = np.zeros((self.n_slots, self._n_k, 9))
        n = C.c_int64(0)
        self._ck(self._L.csmc_pt_get_ssf(self._h, _p(sums), C.byref(n)))
        return np.ascontiguousarray(np.transpose(sums, (0, 2, 1))), n.value

    def pt_stats(self):
        a = np.zeros(self.n_slots); e = np.zeros(self.n_slots)
        self._ck(self._L.csmc_pt_get_stats(self._h, _p(a), _p(e)))
        return a, e
