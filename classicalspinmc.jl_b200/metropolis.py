"""Metropolis variants — mirror of the ``alg`` factories of src/metropolis.jl:181-199.

Each factory returns a callable ``alg(mc, T) -> accepted`` (the reference's
``FunctionWrapper{Float64,Tuple{MonteCarlo,Float64}}``) that performs one sweep on the GPU.  The
drivers recognise these objects by ``kind`` and fuse whole sweep schedules on the device instead of
calling back per sweep; a user-supplied Python callable is still honoured (called once per sweep).

The device sweep visits every site once in colour order instead of N random sites with replacement
(src/metropolis.jl:70) and computes dE from one field evaluation instead of two energy() calls
(:95,98); the Markov chain samples the same Boltzmann distribution (statistical parity,
tests/test_gpu_statistics.py).
"""
from __future__ import annotations


class SweepAlgorithm:
    def __init__(self, kind: str):
        self.kind = kind

    def __call__(self, mc, T: float) -> float:
        """``alg(mc, T)``: one sweep that mutates ``mc.lattice.spins`` like the reference's metropolis!(mc, T).  Called
        on its own it uploads the host spins, sweeps and downloads the result; inside a driver that keeps the state on
        the device (``mc._device_resident``) the copies are skipped."""
        standalone = not getattr(mc, "_device_resident", False)
        eng = mc._upload() if standalone else mc._device()
        if self.kind == "metropolis":
            acc = eng.metropolis(T, 1)
        elif self.kind == "adaptive":
            acc, sig = eng.metropolis_cone(T, mc.sigma, adapt=True, n_sweeps=1)
            mc.sigma = float(sig[0])
        elif self.kind == "fixed_cone":
            acc, _ = eng.metropolis_cone(T, mc.sigma, adapt=False, n_sweeps=1)
        else:
            raise NotImplementedError(
                "MetropolisConstraint / MetropolisConstraintAdaptive are out of scope: they call a global "
                "user constraint per proposal, and metropolis_constraint! is broken in the reference "
                "(undefined e_diff, src/metropolis.jl:44)")
        if standalone:
            mc._download()
        return float(acc[0])


def Metropolis():
    """src/metropolis.jl:181-183"""
    return SweepAlgorithm("metropolis")


def MetropolisAdaptive():
    """src/metropolis.jl:185-187"""
    return SweepAlgorithm("adaptive")


def MetropolisFixedCone():
    """src/metropolis.jl:189-191"""
    return SweepAlgorithm("fixed_cone")


def MetropolisConstraint():
    """src/metropolis.jl:193-195 (unsupported, see SweepAlgorithm.__call__)"""
    return SweepAlgorithm("constraint")


def MetropolisConstraintAdaptive():
    """src/metropolis.jl:197-199 (unsupported)"""
    return SweepAlgorithm("constraint_adaptive")
