"""Model builders shared by the tests: the reference's own test/example Hamiltonians and the
BASELINE.json configurations."""
import math

import numpy as np

import classicalspinmc.jl_b200 as csm


def square_heisenberg(J=-1.0, h=(0.0, 0.0, 0.1)):
    """README.md:35-60 / test/latticetests.jl:21-27 — C1/C2."""
    uc = csm.Square()
    Jm = J * np.eye(3)
    for off in ((1, 0), (-1, 0), (0, 1), (0, -1)):
        csm.addBilinear(uc, 1, 1, Jm, off)
    if h is not None:
        csm.addZeemanCoupling(uc, 1, np.array(h, dtype=float))
    return uc


def kitaev_honeycomb(K=-1.0, G=0.2, Gp=-0.02, J=0.0, h=0.1, J3=0.0):
    """test/mctests.jl:3-33 (J3 as examples/simulated_annealing/honeycomb.jl:66-74) — C3."""
    uc = csm.Honeycomb()
    Jx = np.array([[K + J, Gp, Gp], [Gp, J, G], [Gp, G, J]])
    Jy = np.array([[J, Gp, G], [Gp, K + J, Gp], [G, Gp, J]])
    Jz = np.array([[J, G, Gp], [G, J, Gp], [Gp, Gp, K + J]])
    csm.addBilinear(uc, 1, 2, Jx, (0, -1))
    csm.addBilinear(uc, 1, 2, Jy, (1, -1))
    csm.addBilinear(uc, 1, 2, Jz, (0, 0))
    if J3 != 0.0:
        J3m = J3 * np.eye(3)
        csm.addBilinear(uc, 1, 2, J3m, (1, 0))
        csm.addBilinear(uc, 1, 2, J3m, (1, -2))
        csm.addBilinear(uc, 1, 2, J3m, (-1, 0))
    field = h * np.array([1.0, 1.0, 1.0]) / math.sqrt(3)
    csm.addZeemanCoupling(uc, 1, field)
    csm.addZeemanCoupling(uc, 2, field)
    return uc


def pyrochlore_local(Jxx=0.043, Jyy=0.065, Jzz=0.043, B=1.0):
    """examples/parallel_tempering/pyrochlore.jl:60-86, input_file.jl:36-50, runner.jl:27-30 — C4."""
    uc = csm.Pyrochlore()
    J = np.diag([Jxx, Jyy, Jzz])
    for b1, b2 in ((1, 2), (1, 3), (1, 4), (2, 3), (2, 4), (3, 4)):
        csm.addBilinear(uc, b1, b2, J, (0, 0, 0))
    csm.addBilinear(uc, 1, 2, J, (1, 0, 0))
    csm.addBilinear(uc, 1, 3, J, (0, 1, 0))
    csm.addBilinear(uc, 1, 4, J, (0, 0, 1))
    csm.addBilinear(uc, 2, 3, J, (-1, 1, 0))
    csm.addBilinear(uc, 2, 4, J, (-1, 0, 1))
    csm.addBilinear(uc, 3, 4, J, (0, 1, -1))
    k_B = 1 / 11.6
    mu_B = 0.67 * k_B
    g = np.array([0.0, 0.0, 2.18])
    hdir = np.array([1.0, 0.0, 0.0])
    zs = [np.array(z) / math.sqrt(3) for z in ([1, 1, 1], [1, -1, -1], [-1, 1, -1], [-1, -1, 1])]
    for b, z in enumerate(zs, start=1):
        csm.addZeemanCoupling(uc, b, (hdir @ z) * g * B * mu_B)
    return uc


def triangular_multispin(J=1.0, cubic_scale=0.1, quartic_scale=0.05, onsite=None):
    """NN Heisenberg + up-triangle cubic (3 perspectives) + rhombus
    quartic (4 perspectives), fixed random tensors (seeds 7, 8)."""
    uc = csm.Triangular()
    Jm = J * np.eye(3)
    for off in ((1, 0), (-1, 0), (0, 1), (0, -1), (1, -1), (-1, 1)):
        csm.addBilinear(uc, 1, 1, Jm, off)
    C3 = np.random.default_rng(7).uniform(-cubic_scale, cubic_scale, (3, 3, 3))
    # sites (i, i+a1, i+a2); the same physical term seen from each of its members
    csm.addCubic(uc, 1, 1, 1, C3, (1, 0), (0, 1))
    csm.addCubic(uc, 1, 1, 1, np.transpose(C3, (1, 0, 2)), (-1, 0), (-1, 1))
    csm.addCubic(uc, 1, 1, 1, np.transpose(C3, (2, 1, 0)), (1, -1), (0, -1))
    R4 = np.random.default_rng(8).uniform(-quartic_scale, quartic_scale, (3, 3, 3, 3))
    # sites (i, i+a1, i+a2, i+a1+a2)
    csm.addQuartic(uc, 1, 1, 1, 1, R4, (1, 0), (0, 1), (1, 1))
    csm.addQuartic(uc, 1, 1, 1, 1, np.transpose(R4, (1, 0, 2, 3)), (-1, 0), (-1, 1), (0, 1))
    csm.addQuartic(uc, 1, 1, 1, 1, np.transpose(R4, (2, 1, 0, 3)), (1, -1), (0, -1), (1, 0))
    csm.addQuartic(uc, 1, 1, 1, 1, np.transpose(R4, (3, 1, 2, 0)), (0, -1), (-1, 0), (-1, -1))
    if onsite is not None:
        csm.addOnSite(uc, 1, np.asarray(onsite, dtype=float))
    return uc


def mixed_basis_multispin():
    """A 2-basis 2-D model with cubic and quartic terms across sublattices plus on-site anisotropy:
    exercises every perspective branch of src/lattice.jl:209-283."""
    rng = np.random.default_rng(11)
    uc = csm.Honeycomb()
    csm.addBilinear(uc, 1, 2, rng.normal(size=(3, 3)), (0, 0))
    csm.addBilinear(uc, 1, 2, rng.normal(size=(3, 3)), (0, -1))
    csm.addBilinear(uc, 2, 1, rng.normal(size=(3, 3)), (1, 0))
    csm.addBilinear(uc, 1, 1, rng.normal(size=(3, 3)), (1, 0))
    csm.addBilinear(uc, 1, 1, rng.normal(size=(3, 3)), (-1, 0))
    csm.addCubic(uc, 1, 2, 1, 0.3 * rng.normal(size=(3, 3, 3)), (0, 0), (1, 0))
    csm.addCubic(uc, 2, 1, 2, 0.3 * rng.normal(size=(3, 3, 3)), (0, 1), (1, 0))
    csm.addQuartic(uc, 1, 2, 1, 2, 0.2 * rng.normal(size=(3, 3, 3, 3)), (0, 0), (1, 0), (1, 0))
    csm.addOnSite(uc, 1, np.diag([0.0, 0.0, -0.3]))
    csm.addOnSite(uc, 2, np.array([[0.1, 0.02, 0.0], [0.02, -0.1, 0.0], [0.0, 0.0, 0.05]]))
    csm.addZeemanCoupling(uc, 1, np.array([0.1, -0.2, 0.3]))
    csm.addZeemanCoupling(uc, 2, np.array([-0.3, 0.1, 0.2]))
    return uc


def chain_heisenberg(J=-1.0):
    """Open 1-D Heisenberg chain (Fisher's exact solution) — analytic known answer."""
    uc = csm.UnitCell(np.array([1.0]))
    csm.addBasisSite(uc, np.array([0.0]))
    Jm = J * np.eye(3)
    csm.addBilinear(uc, 1, 1, Jm, (1,))
    csm.addBilinear(uc, 1, 1, Jm, (-1,))
    return uc
