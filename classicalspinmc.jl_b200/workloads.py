"""The BASELINE.json workloads C2-C5: the reference's own test / example Hamiltonians at
the benchmark sizes.  bench.py, __graft_entry__.py, tools/ and the tests all take their models from here."""
import math
import types

import numpy as np

from . import bravais as _bravais
from . import unit_cell as _unit_cell

csm = types.SimpleNamespace(Square=_bravais.Square, Honeycomb=_bravais.Honeycomb, Pyrochlore=_bravais.Pyrochlore,
                            Triangular=_bravais.Triangular, addBilinear=_unit_cell.addBilinear,
                            addZeemanCoupling=_unit_cell.addZeemanCoupling, addCubic=_unit_cell.addCubic,
                            addQuartic=_unit_cell.addQuartic, addOnSite=_unit_cell.addOnSite)


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



# parallel-tempering temperature grids: examples/parallel_tempering/runner.jl:14 (log-spaced),
# input_file.jl:55-56 ([0.09, 14] k_B with k_B = 1/11.6) for C4; [0.01, 1.0] for C3
PT_DEFAULTS = {"C3": dict(R=64, Tmin=0.01, Tmax=1.0), "C4": dict(R=128, Tmin=0.09 / 11.6, Tmax=14 / 11.6)}


def unit_cell(name):
    return {"C2": lambda: square_heisenberg(J=-1.0, h=(0.0, 0.0, 0.1)), "C3": kitaev_honeycomb, "C4": pyrochlore_local,
            "C5": triangular_multispin}[name]()


def workload_model(name, L=None):
    """(ModelData, config dict) of BASELINE.json configs[1..4] ("C2".."C5"), optionally at another linear size."""
    from ._abi import ModelData
    if name == "C2":
        L = L or 1024
        return ModelData(unit_cell(name), (L, L), 1.0), dict(
            workload="C2", lattice="square", L=L, model="Heisenberg J=-1 + h_z=0.1", T=1.0, replicas_per_gpu=1)
    if name == "C3":
        L = L or 256
        return ModelData(unit_cell(name), (L, L), 1.0), dict(
            workload="C3", lattice="honeycomb", L=L, model="Kitaev-Gamma K=-1 G=0.2 Gp=-0.02 h=0.1[111]")
    if name == "C4":
        L = L or 32
        return ModelData(unit_cell(name), (L, L, L), 0.5), dict(
            workload="C4", lattice="pyrochlore", L=L, model="local-frame Jxx/Jyy/Jzz + Zeeman")
    if name == "C5":
        L = L or 512
        return ModelData(unit_cell(name), (L, L), 1.0), dict(
            workload="C5", lattice="triangular", L=L, model="Heisenberg + cubic + quartic")
    raise ValueError(f"unknown workload {name}")
