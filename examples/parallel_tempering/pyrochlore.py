"""Model helpers for the pyrochlore example (the reference ships the same helpers with its example,
examples/parallel_tempering/pyrochlore.jl): the 12 nearest-neighbour bonds of the four-site unit cell."""
import numpy as np

import classicalspinmc.jl_b200 as csm

PAIRS = ((1, 2), (1, 3), (1, 4), (2, 3), (2, 4), (3, 4))
# bond within the unit cell ("A" tetrahedron) and to the neighbouring cell ("B" tetrahedron)
B_OFFSETS = ((1, 0, 0), (0, 1, 0), (0, 0, 1), (-1, 1, 0), (-1, 0, 1), (0, 1, -1))


def _add(uc, matrices):
    for (b1, b2), J in zip(PAIRS, matrices):
        csm.addBilinear(uc, b1, b2, J, (0, 0, 0))
    for (b1, b2), J, off in zip(PAIRS, matrices, B_OFFSETS):
        csm.addBilinear(uc, b1, b2, J, off)


def addInteractionsLocal(uc, params):
    """diag(Jxx, Jyy, Jzz) on every bond, spins expressed in their local frames."""
    J = np.diag([float(params.get(k, 0.0)) for k in ("Jxx", "Jyy", "Jzz")])
    _add(uc, [J] * 6)


def addInteractionsGlobal(uc, params):
    """The same model with the spins in the global cubic frame: one 3x3 matrix per sublattice pair."""
    Jx, Jy, Jz = (float(params.get(k, 0.0)) for k in ("Jxx", "Jyy", "Jzz"))
    J1 = (-2 * Jx + 2 * Jz) / 6
    J2 = (-2 * Jx - Jz) / 3
    J3 = (Jx - 3 * Jy + 2 * Jz) / 6
    J4 = (-Jx - 3 * Jy - 2 * Jz) / 6
    M = {
        (1, 2): [[J1, J2, -J1], [J3, -J1, J4], [-J4, -J1, -J3]],
        (1, 3): [[-J1, J1, J2], [J4, J3, -J1], [-J3, -J4, -J1]],
        (1, 4): [[-J1, -J1, -J2], [J4, -J3, J1], [-J3, J4, J1]],
        (2, 3): [[-J3, -J4, -J1], [J1, -J1, -J2], [-J4, -J3, J1]],
        (2, 4): [[-J3, J4, J1], [J1, J1, J2], [-J4, J3, -J1]],
        (3, 4): [[-J4, J3, -J1], [-J3, J4, J1], [J1, J1, J2]],
    }
    _add(uc, [np.array(M[p], dtype=float) for p in PAIRS])
