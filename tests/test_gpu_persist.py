"""The tile-resident persistent kernel (csmc_persist, CSMC_FLAG_PERSIST) against the per-colour pass kernels: the same
sequences of sweeps must leave bit-identical spins and acceptance counts (same per-site arithmetic and Philox counters;
only the place the neighbours are read from changes: shared-memory tiles + halos exchanged through L2 instead of global
memory).  Reference semantics of the updates: src/monte_carlo.jl:126-139,201-213, src/metropolis.jl:65-101."""
import os

import numpy as np
import pytest

from classicalspinmc.jl_b200 import _lib
from classicalspinmc.jl_b200._abi import (FLAG_JIT, FLAG_NO_AUTOTUNE, FLAG_NO_PERSIST, FLAG_NO_RESIDENT, FLAG_PERSIST,
                                          ModelData)
from oracle import oracle as orc
from tests import models

pytestmark = pytest.mark.gpu

BASE = FLAG_JIT | FLAG_NO_RESIDENT | FLAG_NO_AUTOTUNE

CASES = [
    # name, builder, shape, S, replicas, forced tile grid (None: planner's choice)
    ("square-192", lambda: models.square_heisenberg(), (192, 192), 1.0, 1, None),
    ("square-192-ragged-5x7", lambda: models.square_heisenberg(), (192, 192), 1.0, 1, "5x7"),
    ("square-192-strips-9x1", lambda: models.square_heisenberg(), (192, 192), 1.0, 1, "9x1"),
    ("square-192-columns-1x6", lambda: models.square_heisenberg(), (192, 192), 1.0, 1, "1x6"),
    ("square-192-3x3", lambda: models.square_heisenberg(), (192, 192), 1.0, 1, "3x3"),
    ("square-128-3-replicas", lambda: models.square_heisenberg(), (128, 128), 1.0, 3, None),
    ("honeycomb-J3-128x96-5-replicas", lambda: models.kitaev_honeycomb(J3=0.25), (128, 96), 1.0, 5, None),
    ("pyrochlore-16x12x24-2-replicas", lambda: models.pyrochlore_local(), (16, 12, 24), 0.5, 2, None),
    ("triangular-multispin-onsite-192x176", lambda: models.triangular_multispin(onsite=np.diag([0.1, -0.2, 0.3])), (192, 176), 1.0, 1, None),
    ("mixed-basis-multispin-160x144", lambda: models.mixed_basis_multispin(), (160, 144), 0.8, 2, None),
]


def _run(eng, R):
    T = np.geomspace(0.3, 2.0, R)
    eng.randomize(4242)
    eng.set_temperatures(T)
    eng.cycles_async(3, 4, 1)            # graph replay: 4 OR + 1 Metropolis, three times
    eng.sync()
    eng.overrelax(7)
    acc = eng.metropolis(T, 3)
    acc2, _ = eng.metropolis_cone(T, 0.4, adapt=False, n_sweeps=2)
    eng.deterministic(2)
    return [eng.get_spins(r) for r in range(R)], acc, acc2, eng.total_energy()


@pytest.mark.parametrize("name,builder,shape,S,R,grid", CASES, ids=[c[0] for c in CASES])
def test_persistent_kernel_is_bit_identical_to_pass_kernels(name, builder, shape, S, R, grid, monkeypatch):
    md = ModelData(builder(), shape, S)
    ref = _lib.Engine(md, n_replicas=R, seed=77, flags=BASE | FLAG_NO_PERSIST)
    assert ref.persist_info()[0] == 0
    want = _run(ref, R)
    ref.close()
    if grid:
        monkeypatch.setenv("CSMC_PERSIST_GRID", grid)
    eng = _lib.Engine(md, n_replicas=R, seed=77, flags=BASE | FLAG_PERSIST)
    tiles, g, nrep, smem, _ = eng.persist_info()
    assert tiles > 0, "the tile-resident kernel did not load"
    if grid:
        assert f"{g[0]}x{g[1]}" == grid
    got = _run(eng, R)
    l0 = eng.launches
    eng.overrelax(7)
    launches = eng.launches - l0
    for r in range(R):
        assert np.array_equal(got[0][r], want[0][r]), f"replica {r}"
    assert np.array_equal(got[1], want[1]) and np.array_equal(got[2], want[2]) and np.array_equal(got[3], want[3])
    # and it really ran on the persistent kernel: one launch per batch of replicas instead of one per colour pass
    assert launches == -(-R // nrep), launches


def test_persistent_kernel_against_oracle():
    """... and directly against the oracle in colour order (overrelaxation, same-stream Metropolis, deterministic)."""
    seed = 99
    md = ModelData(models.kitaev_honeycomb(J3=0.25), (160, 128), 1.0)
    lat = orc.OracleLattice(md)
    eng = _lib.Engine(md, seed=seed, flags=BASE | FLAG_PERSIST)
    assert eng.persist_info()[0] > 0
    s = lat.randomize(seed=5)
    eng.set_spins(s)
    order = eng.colour_order()
    kappa = np.zeros(lat.N)
    eng.overrelax(2)
    for _ in range(2):
        lat.sweep_tracked(s, order, 0, kappa)
    assert np.all(np.abs(eng.get_spins() - s).max(axis=1) <= 1e-12 * np.maximum(kappa, 1.0))
    eng.set_spins(s)
    acc = eng.metropolis(0.6, 2)[0]
    acc_ref = lat.metropolis_philox(s, order, 0.6, seed, 0, 0) + lat.metropolis_philox(s, order, 0.6, seed, 0, 1)
    assert acc == acc_ref
    assert np.abs(eng.get_spins() - s).max() <= 1e-12
    eng.deterministic(2)
    lat.deterministic(s, order, 2)
    assert np.abs(eng.get_spins() - s).max() <= 1e-10


def test_persistent_kernel_in_parallel_tempering_is_bit_identical():
    """The PT loop (OR block + Metropolis sweep as one sequence per Metropolis block) on the persistent kernel."""
    md = ModelData(models.kitaev_honeycomb(), (96, 96), 1.0)
    R = 6
    T = np.geomspace(0.2, 1.5, R)
    p = dict(t_thermalization=40, t_measurement=80, probe_rate=10, swap_rate=10, overrelaxation_rate=5)
    out = []
    for flags in (BASE | FLAG_NO_PERSIST, BASE | FLAG_PERSIST):
        eng = _lib.Engine(md, n_replicas=R, seed=5, flags=flags)
        eng.randomize(31)
        eng.pt_init(T)
        eng.pt_run(p, 0, 120)
        out.append((eng.pt_series(), eng.pt_slots(), eng.pt_stats(), [eng.get_spins(r) for r in range(R)], eng.persist_info()[0]))
        eng.close()
    a, b = out
    assert a[4] == 0 and b[4] > 0
    assert np.array_equal(a[0][0], b[0][0]) and np.array_equal(a[0][1], b[0][1]) and np.array_equal(a[1], b[1])
    assert np.array_equal(a[2][0], b[2][0]) and np.array_equal(a[2][1], b[2][1])
    assert all(np.array_equal(x, y) for x, y in zip(a[3], b[3]))


def test_generator_reports_flops_per_update():
    """csmc_kernel_costs: the code generator's own count of fp64 flops per overrelaxation update (fma = 2, a literal
    +-1 coefficient = 1) feeds the fp64 roofline of bench.py.  Square Heisenberg J = -1: 4 neighbours x 3 additions + 21
    for F = g - h, s.F, F.F, the division and the reflection (src/monte_carlo.jl:126-139)."""
    eng = _lib.Engine(ModelData(models.square_heisenberg(), (192, 192), 1.0), flags=BASE | FLAG_NO_PERSIST)
    flops, nbytes = eng.kernel_costs()
    assert flops == 33.0 and nbytes == 72.0
    eng = _lib.Engine(ModelData(models.triangular_multispin(), (192, 176), 1.0), flags=BASE | FLAG_NO_PERSIST)
    flops, nbytes = eng.kernel_costs()
    # 4 quartic slots x 3 x (27 + 9 + 3) fma + 3 cubic slots x 3 x (9 + 3) fma + 6 x 3 additions + 21, minus the first
    # multiply of each chain: about a thousand flops per site
    assert 900.0 < flops < 1100.0 and nbytes == 120.0
