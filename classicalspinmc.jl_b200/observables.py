"""Observables — mirror of src/observables.jl.

``ErrorPropagator`` restates the parts of BinningAnalysis.jl 0.6.1 (Manifest.toml:29-33, not vendored
in the reference) that src/observables.jl:32-63 and src/hdf5.jl:220-227 call: logarithmic binning of
N-argument samples with per-level sums and second-moment matrices, ``mean``, ``var`` with a gradient
(first-order error propagation), ``std_error`` at the "reliable level".  PARITY UNPINNED: no reference
test calls specific_heat / susceptibility / write_final_observables and the package source is not in
/root/reference; means are exact by construction, error bars follow the published algorithm.
"""
from __future__ import annotations

import math

import numpy as np

N_LEVELS = 32


class ErrorPropagator:
    """Log-binning accumulator for vectors of `n_args` observables (ErrorPropagator{Float64,32})."""

    def __init__(self, n_args: int = 2):
        self.n_args = n_args
        self.sums1D = np.zeros((N_LEVELS, n_args))
        self.sums2D = np.zeros((N_LEVELS, n_args, n_args))
        self.count = np.zeros(N_LEVELS, dtype=np.int64)
        self._held = np.zeros((N_LEVELS, n_args))
        self._full = np.zeros(N_LEVELS, dtype=bool)

    def push(self, *args):
        x = np.asarray(args, dtype=np.float64)
        lvl = 0
        while lvl < N_LEVELS:
            self.sums1D[lvl] += x
            self.sums2D[lvl] += np.outer(x, x)
            self.count[lvl] += 1
            if not self._full[lvl]:
                self._held[lvl] = x
                self._full[lvl] = True
                return
            self._full[lvl] = False
            x = 0.5 * (self._held[lvl] + x)
            lvl += 1

    def __len__(self):
        return int(self.count[0])

    def reliable_level(self) -> int:
        """highest binning level that still has at least 32 bins (0-based), else 0."""
        ok = np.nonzero(self.count >= 32)[0]
        return int(ok[-1]) if len(ok) else 0

    def means(self, lvl: int = 0):
        return self.sums1D[lvl] / max(self.count[lvl], 1)

    def mean(self, arg=None, lvl: int = 0):
        """mean(ep, i) (1-based argument index, as the reference calls it) or mean(ep, f) = f(means)."""
        if callable(arg):
            return arg(self.means(lvl))
        if arg is None:
            return self.means(lvl)
        return float(self.means(lvl)[arg - 1])

    def covmat(self, lvl: int):
        n = self.count[lvl]
        if n < 2:
            return np.full((self.n_args, self.n_args), np.nan)
        m = self.sums1D[lvl]
        return (self.sums2D[lvl] - np.outer(m, m) / n) / (n - 1)

    def var(self, gradient, lvl=None):
        lvl = self.reliable_level() if lvl is None else lvl
        g = np.asarray(gradient(self.means(0)) if callable(gradient) else gradient, dtype=np.float64)
        return float(g @ self.covmat(lvl) @ g)

    def std_error(self, arg: int = 1, lvl=None):
        lvl = self.reliable_level() if lvl is None else lvl
        n = self.count[lvl]
        if n < 2:
            return float("nan")
        return math.sqrt(abs(self.covmat(lvl)[arg - 1, arg - 1]) / n)


class Observables:
    """src/observables.jl:5-10 (the optional structure-factor binner is out of scope)."""

    def __init__(self, N_k: int = 0):
        self.energy = ErrorPropagator(2)
        self.magnetization = ErrorPropagator(2)
        self.correlations = None


def get_magnetization(lattice) -> float:
    """src/observables.jl:12-18: norm of the vector sum of all spins (computed on the GPU)."""
    lattice.upload()
    m = lattice.engine().magnetization_vector()[0]
    return float(np.sqrt(m @ m))


def update_observables(mc, energy: float, magnetization: float):
    """src/observables.jl:20-25"""
    mc.observables.energy.push(energy, energy ** 2)
    mc.observables.magnetization.push(magnetization, magnetization ** 2)


def std_error_tweak(ep: ErrorPropagator, gradient, lvl=None):
    """src/observables.jl:32-34"""
    lvl = ep.reliable_level() if lvl is None else lvl
    return math.sqrt(abs(ep.var(gradient, lvl) / ep.count[lvl]))


def _specific_heat(ep, temp, n_sites):
    c = lambda e: 1 / temp ** 2 * (e[1] - e[0] * e[0]) / n_sites                    # :42
    dc = lambda e: np.array([-2.0 * 1 / temp ** 2 * e[0] / n_sites, 1 / temp ** 2 / n_sites])  # :43
    return ep.mean(c), std_error_tweak(ep, dc)


def _susceptibility(ep, temp, n_sites):
    x = lambda m: 1 / temp * (m[1] - m[0] * m[0]) / n_sites                         # :57
    dx = lambda m: np.array([-2 * 1 / temp * m[0] / n_sites, 1 / temp / n_sites])   # :58
    return ep.mean(x), std_error_tweak(ep, dx)


def specific_heat(mc):
    """src/observables.jl:36-49 -> (c, dc)"""
    return _specific_heat(mc.observables.energy, mc.T, mc.lattice.size)


def susceptibility(mc):
    """src/observables.jl:51-63 -> (chi, dchi)"""
    return _susceptibility(mc.observables.magnetization, mc.T, mc.lattice.size)
