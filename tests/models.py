"""Model builders shared by the tests: the BASELINE workloads (classicalspinmc.jl_b200.workloads) plus
test-only Hamiltonians that exercise every perspective branch and the analytic known answers."""
import numpy as np

import classicalspinmc.jl_b200 as csm
from classicalspinmc.jl_b200.workloads import (kitaev_honeycomb, pyrochlore_local, square_heisenberg,  # noqa: F401
                                               triangular_multispin)


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
