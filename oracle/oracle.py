"""ctypes binding of oracle/liboracle.so (CPU restatement of the reference).

TEST INFRASTRUCTURE ONLY — import from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs, never from the product package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_FAST = None
FAST_CFLAGS = "-O3 -march=native -ffp-contract=fast -fno-math-errno -fPIC -fopenmp -std=gnu11"

f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
i64p = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "csmc_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return so


def build_fast() -> str:
    """Performance build of the same source for the TIMED CPU legs of bench.py (cpu_baseline, --impl reference):
    -O3 -march=native with FMA contraction allowed.  It is not the bit-exact checker (the parity tests use the
    -ffp-contract=off build above).  -march=native is only valid on the machine that compiled it, so the library is
    rebuilt on the box that runs it, into a directory keyed by the host's CPU flags."""
    import hashlib
    import platform
    try:
        flags = next(l for l in open("/proc/cpuinfo") if l.startswith("flags"))
    except Exception:
        flags = platform.processor()
    key = hashlib.sha1((flags + FAST_CFLAGS).encode()).hexdigest()[:12]
    out_dir = os.path.join(_HERE, "_fast")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, f"liboracle_fast_{key}.so")
    src = os.path.join(_HERE, "csmc_oracle.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        cc = "/usr/bin/gcc" if os.access("/usr/bin/gcc", os.X_OK) else "gcc"
        tmp = so + f".{os.getpid()}.tmp"
        subprocess.check_call([cc] + FAST_CFLAGS.split() + ["-shared", "-o", tmp, src, "-lm"])
        os.replace(tmp, so)
    return so


def lib(fast: bool = False):
    global _LIB, _FAST
    if fast:
        if _FAST is None:
            _FAST = _declare(C.CDLL(build_fast()))
        return _FAST
    if _LIB is not None:
        return _LIB
    _LIB = _declare(C.CDLL(build()))
    return _LIB


def _declare(L):
    vp = C.c_void_p
    L.orc_build.restype = vp
    L.orc_build.argtypes = [vp, C.c_int]
    L.orc_free.argtypes = [vp]
    L.orc_n_sites.restype = C.c_int64
    L.orc_n_sites.argtypes = [vp]
    L.orc_get_tables.argtypes = [vp, vp, vp, vp]
    L.orc_get_bilinear_matrices.argtypes = [vp, f64p]
    L.orc_local_field.argtypes = [vp, f64p, C.c_int64, f64p]
    L.orc_site_energy.restype = C.c_double
    L.orc_site_energy.argtypes = [vp, f64p, C.c_int64]
    L.orc_total_energy.restype = C.c_double
    L.orc_total_energy.argtypes = [vp, f64p, C.POINTER(C.c_double)]
    L.orc_magnetization.restype = C.c_double
    L.orc_magnetization.argtypes = [vp, f64p, vp]
    L.orc_overrelax.argtypes = [vp, f64p, vp, C.c_int64, C.c_int]
    L.orc_deterministic_order.argtypes = [vp, f64p, vp, C.c_int64, C.c_int]
    L.orc_randomize_spins.argtypes = [vp, f64p, C.c_uint64, C.c_uint32]
    L.orc_metropolis_philox.restype = C.c_double
    L.orc_metropolis_philox.argtypes = [vp, f64p, vp, C.c_int64, C.c_double, C.c_double,
                                        C.c_uint64, C.c_uint32, C.c_uint64]
    L.orc_adapt_sigma.restype = C.c_double
    L.orc_adapt_sigma.argtypes = [C.c_double, C.c_double, C.c_double]
    L.orc_simulated_annealing.argtypes = [vp, f64p, f64p, C.c_int, C.c_int64, C.c_int, C.c_int,
                                          C.c_double, C.c_uint64, vp]
    L.orc_deterministic_updates.argtypes = [vp, f64p, C.c_int64, C.c_uint64]
    L.orc_exchange_accept.restype = C.c_int
    L.orc_exchange_accept.argtypes = [C.c_double] * 5
    L.orc_parallel_tempering.restype = C.c_int64
    L.orc_parallel_tempering.argtypes = [vp, f64p, f64p, C.c_int, C.c_int64, C.c_int64, C.c_int,
                                         C.c_int, C.c_int, C.c_uint64, C.c_int, vp, vp, vp, vp]
    L.orc_structure_factor.argtypes = [vp, f64p, f64p, f64p, C.c_int64, f64p]
    L.orc_cycles.restype = C.c_double
    L.orc_cycles.argtypes = [vp, f64p, C.c_int, C.c_double, C.c_int64, C.c_int, C.c_int, C.c_uint64]
    L.orc_max_threads.restype = C.c_int
    L.orc_philox_raw.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint32, vp]
    L.orc_local_field_all.argtypes = [vp, f64p, f64p]
    L.orc_sweep_tracked.restype = C.c_double
    L.orc_sweep_tracked.argtypes = [vp, f64p, vp, C.c_int64, C.c_int, C.c_double, C.c_uint64, C.c_uint32, C.c_uint64, f64p]
    return L


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class OracleLattice:
    """Reference-layout lattice tables built from a ``ModelData`` (classicalspinmc.jl_b200._abi).

    Spin arrays are (N, 3) C-contiguous float64 (== Julia's 3 x N column-major)."""

    def __init__(self, model_data, literal: bool = False, fast: bool = False):
        self._md = model_data
        self._L = lib(fast)
        self._h = self._L.orc_build(C.byref(model_data.struct), 1 if literal else 0)
        self.N = int(self._L.orc_n_sites(self._h))
        self.N2, self.N3, self.N4 = model_data.n2, model_data.n3, model_data.n4
        self.S = model_data.S

    def __del__(self):
        try:
            if self._h:
                self._L.orc_free(self._h)
                self._h = None
        except Exception:
            pass

    # -- tables -------------------------------------------------------------------------------
    def tables(self):
        bil = np.zeros((self.N, self.N2), np.int64)
        cub = np.zeros((self.N, self.N3, 2), np.int64)
        quar = np.zeros((self.N, self.N4, 3), np.int64)
        self._L.orc_get_tables(self._h, _ptr(bil), _ptr(cub), _ptr(quar))
        return bil, cub, quar

    def bilinear_matrices(self):
        out = np.zeros((self.N, self.N2, 9))
        if self.N2:
            self._L.orc_get_bilinear_matrices(self._h, out)
        return out

    # -- Hamiltonian --------------------------------------------------------------------------
    def local_field(self, spins, p):
        out = np.zeros(3)
        self._L.orc_local_field(self._h, spins, p, out)
        return out

    def local_field_all(self, spins):
        out = np.zeros((self.N, 3))
        self._L.orc_local_field_all(self._h, spins, out)
        return out

    def site_energy(self, spins, p):
        return self._L.orc_site_energy(self._h, spins, p)

    def site_energy_all(self, spins):
        return np.array([self._L.orc_site_energy(self._h, spins, p) for p in range(1, self.N + 1)])

    def total_energy(self, spins, with_abs=False):
        a = C.c_double(0.0)
        e = self._L.orc_total_energy(self._h, spins, C.byref(a))
        return (e, a.value) if with_abs else e

    def magnetization(self, spins, vector=False):
        m3 = np.zeros(3)
        m = self._L.orc_magnetization(self._h, spins, _ptr(m3))
        return m3 if vector else m

    # -- updates ------------------------------------------------------------------------------
    def overrelax(self, spins, order=None, n_sweeps=1):
        o = None if order is None else np.ascontiguousarray(order, np.int64)
        self._L.orc_overrelax(self._h, spins, _ptr(o), 0 if o is None else len(o), n_sweeps)

    def deterministic(self, spins, order=None, n_sweeps=1):
        o = None if order is None else np.ascontiguousarray(order, np.int64)
        self._L.orc_deterministic_order(self._h, spins, _ptr(o), 0 if o is None else len(o), n_sweeps)

    def sweep_tracked(self, spins, order, kind, kappa, T=1.0, seed=0, replica=0, sweep_ctr=0):
        """One colour-order sweep (kind 0 overrelaxation, 1 deterministic, 2 same-stream Metropolis) that also advances
        the per-site forward error bound ``kappa`` (units of TOL * S, see csmc_oracle.c); returns the accepted count."""
        o = np.ascontiguousarray(order, np.int64)
        return self._L.orc_sweep_tracked(self._h, spins, _ptr(o), len(o), kind, T, seed, replica, sweep_ctr, kappa)

    def randomize(self, seed, replica=0):
        s = np.zeros((self.N, 3))
        self._L.orc_randomize_spins(self._h, s, seed, replica)
        return s

    def metropolis_philox(self, spins, order, T, seed, replica, sweep_ctr, sigma=-1.0):
        o = None if order is None else np.ascontiguousarray(order, np.int64)
        return self._L.orc_metropolis_philox(self._h, spins, _ptr(o), self.N if o is None else len(o),
                                             T, sigma, seed, replica, sweep_ctr)

    def simulated_annealing(self, spins, temps, t_thermalization, rate, alg=0, sigma0=60.0, seed=1):
        temps = np.ascontiguousarray(temps, np.float64)
        acc = np.zeros(len(temps))
        self._L.orc_simulated_annealing(self._h, spins, temps, len(temps), t_thermalization, rate,
                                        alg, sigma0, seed, _ptr(acc))
        return acc

    def deterministic_updates(self, spins, t_deterministic, seed=2):
        self._L.orc_deterministic_updates(self._h, spins, t_deterministic, seed)

    def parallel_tempering(self, spins, T, t_th, t_meas, probe_rate, swap_rate, rate, seed=3, n_threads=0):
        T = np.ascontiguousarray(T, np.float64)
        R = len(T)
        n_max = max(1, t_meas // max(probe_rate, 1) + 2)
        E = np.zeros((n_max, R)); M = np.zeros((n_max, R))
        acc = np.zeros(R); exch = np.zeros(R)
        n = self._L.orc_parallel_tempering(self._h, spins, T, R, t_th, t_meas, probe_rate, swap_rate,
                                           rate, seed, n_threads, _ptr(E), _ptr(M), _ptr(acc), _ptr(exch))
        return E[:n], M[:n], acc, exch

    def structure_factor(self, spins, site_positions, ks):
        """site_positions: (D, N) and ks: (D, N_k) as in the reference; returns (9, N_k)."""
        pos = np.ascontiguousarray(np.asarray(site_positions, dtype=np.float64).T)      # column-major D x N
        kk = np.ascontiguousarray(np.asarray(ks, dtype=np.float64).T)
        out = np.zeros((kk.shape[0], 9))
        self._L.orc_structure_factor(self._h, spins, pos, kk, kk.shape[0], out)
        return np.ascontiguousarray(out.T)

    def cycles(self, spins, n_threads, T, n_cycles, or_per_cycle, metro_per_cycle, seed=5):
        return self._L.orc_cycles(self._h, spins, n_threads, T, n_cycles, or_per_cycle, metro_per_cycle, seed)


def annealing_temperatures(T_target, schedule, T0=1.0):
    """Temperature sequence of simulated_annealing! (src/monte_carlo.jl:159,168,184-185)."""
    temps, T, time = [], T0, 1
    while T > T_target:
        temps.append(T)
        T = schedule(time)
        time += 1
    return temps


def exchange_accept(T, E, Tp, Ep, u):
    return bool(lib().orc_exchange_accept(T, E, Tp, Ep, u))


def adapt_sigma(sigma, accepted, n):
    return lib().orc_adapt_sigma(sigma, accepted, n)


def philox(seed, c0, c1, ctr, tag):
    out = np.zeros(4, np.uint32)
    lib().orc_philox_raw(seed, c0, c1, ctr, tag, _ptr(out))
    return out


def max_threads():
    return lib().orc_max_threads()
