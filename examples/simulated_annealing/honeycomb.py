"""Model helpers for the honeycomb examples (the reference ships the same helpers with its example,
examples/simulated_annealing/honeycomb.jl): nearest- and third-neighbour bonds of a two-site unit cell."""
import numpy as np

import classicalspinmc.jl_b200 as csm

NN_BONDS = ((0, -1), (1, -1), (0, 0))          # x, y, z bond: offset of the sublattice-2 site
THIRD_BONDS = ((1, 0), (1, -2), (-1, 0))


def add_bonds(uc, nn_matrices, J3=0.0):
    for J, off in zip(nn_matrices, NN_BONDS):
        csm.addBilinear(uc, 1, 2, np.asarray(J, dtype=float), off)
    for off in THIRD_BONDS:
        csm.addBilinear(uc, 1, 2, J3 * np.eye(3), off)      # dropped by addBilinear when J3 == 0


def addInteractionsKitaev(uc, params):
    """J1-K-Gamma-Gamma' (+J3) model in the cubic Kitaev frame; missing keys are zero."""
    J, K, G, Gp, J3 = (float(params.get(k, 0.0)) for k in ("J1", "K", "G", "Gp", "J3"))

    def bond(axis):
        M = np.full((3, 3), Gp)
        others = [a for a in range(3) if a != axis]
        M[others[0], others[1]] = M[others[1], others[0]] = G
        M[np.diag_indices(3)] = J
        M[axis, axis] += K
        return M

    add_bonds(uc, [bond(0), bond(1), bond(2)], J3)


def addInteractionsCartesian(uc, params):
    """XXZ + bond-dependent anisotropy (D, E) in the global frame; the x and y bonds are the z bond rotated by
    -/+ 120 degrees about the c axis."""
    p = {k: float(params.get(k, 0.0)) for k in ("J1xy", "J1z", "D", "E", "J3xy", "J3z")}
    c, s = -0.5, np.sqrt(3) / 2
    U = np.array([[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]])
    Jz = np.array([[p["J1xy"] + p["D"], p["E"], 0.0], [p["E"], p["J1xy"] + p["D"], 0.0], [0.0, 0.0, p["J1z"]]])
    for J, off in zip((U @ Jz @ U.T, U.T @ Jz @ U, Jz), NN_BONDS):
        csm.addBilinear(uc, 1, 2, J, off)
    J3 = np.diag([p["J3xy"], p["J3xy"], p["J3z"]])
    for off in THIRD_BONDS:
        csm.addBilinear(uc, 1, 2, J3, off)
