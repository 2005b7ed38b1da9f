"""CPU tests (no GPU): libcsmc.so loads and exports every symbol include/csmc.h declares, and the
host-side planning (colouring of the interaction hypergraph, storage permutation) is valid against
the oracle's reference-layout tables."""
import ctypes
import os
import re

import numpy as np
import pytest

from classicalspinmc.jl_b200 import _lib
from classicalspinmc.jl_b200._abi import FLAG_FORCE_GENERIC, ModelData
from oracle import oracle as orc
from tests import models

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "csmc.h")).read()
    declared = set(re.findall(r"\b(csmc_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"csmc_model", "csmc_opts", "csmc_handle", "csmc_pt_params"}
    assert declared == set(_lib.EXPORTS)
    so = ctypes.CDLL(_lib._SO)
    for name in sorted(declared):
        assert hasattr(so, name), name
    assert _lib.lib().csmc_version() == 100


def _conflict_pairs(lat):
    """All unordered pairs of distinct sites that share an interaction term (from oracle tables)."""
    bil, cub, quar = lat.tables()
    pairs = set()
    N = lat.N
    for p in range(N):
        groups = [[p + 1, j] for j in bil[p] if j != 0]
        groups += [[p + 1, c[0], c[1]] for c in cub[p] if c[0] != 0]
        groups += [[p + 1, q[0], q[1], q[2]] for q in quar[p] if q[0] != 0]
        for g in groups:
            for a in g:
                for b in g:
                    if a < b:
                        pairs.add((a - 1, b - 1))
    return pairs


PLAN_CASES = [
    ("square-4x4", lambda: models.square_heisenberg(), (4, 4), "periodic", 2, True),
    ("square-6x8", lambda: models.square_heisenberg(), (6, 8), "periodic", 2, True),
    ("square-2x2", lambda: models.square_heisenberg(), (2, 2), "periodic", 2, True),
    ("square-3x3", lambda: models.square_heisenberg(), (3, 3), "periodic", 3, True),
    ("square-5x7", lambda: models.square_heisenberg(), (5, 7), "periodic", None, None),
    ("square-11x13", lambda: models.square_heisenberg(), (11, 13), "periodic", None, False),
    ("square-open-5x4", lambda: models.square_heisenberg(), (5, 4), "open", 2, True),
    ("honeycomb-4x4", lambda: models.kitaev_honeycomb(), (4, 4), "periodic", 2, True),
    ("honeycomb-J3-5x3", lambda: models.kitaev_honeycomb(J3=0.2), (5, 3), "periodic", 2, True),
    ("pyrochlore-2x3x2", lambda: models.pyrochlore_local(), (2, 3, 2), "periodic", 4, True),
    ("triangular-multispin-4x4", lambda: models.triangular_multispin(), (4, 4), "periodic", 4, True),
    ("triangular-multispin-6x4", lambda: models.triangular_multispin(), (6, 4), "periodic", 4, True),
    ("triangular-multispin-open", lambda: models.triangular_multispin(), (5, 3), "open", 4, True),
    ("mixed-basis-4x4", lambda: models.mixed_basis_multispin(), (4, 4), "periodic", None, True),
    ("mixed-basis-open", lambda: models.mixed_basis_multispin(), (3, 5), "open", None, True),
    ("chain-open-9", lambda: models.chain_heisenberg(), (9,), "open", 2, True),
    ("chain-7", lambda: models.chain_heisenberg(), (7,), "periodic", None, True),
]


@pytest.mark.parametrize("name,builder,shape,bc,ncol,structured", PLAN_CASES, ids=[c[0] for c in PLAN_CASES])
def test_colouring_is_race_free(name, builder, shape, bc, ncol, structured):
    md = ModelData(builder(), shape, 1.0, bc)
    lat = orc.OracleLattice(md)
    for flags in (0, FLAG_FORCE_GENERIC):
        colour, n_colours, st, pos = _lib.plan(md, flags)
        assert colour.min() == 0 and colour.max() == n_colours - 1
        for a, b in _conflict_pairs(lat):
            assert colour[a] != colour[b], f"sites {a},{b} share a term and a colour"
        # storage positions: a permutation into a (padded) colour-major layout
        assert len(set(pos.tolist())) == lat.N
        order = np.argsort(pos)
        assert np.all(np.diff(colour[order]) >= 0), "storage is not colour-major"
        if flags == FLAG_FORCE_GENERIC:
            assert not st
        elif structured is not None:
            assert st == structured
        if ncol is not None:
            assert n_colours == ncol


def test_plan_rejects_bad_models():
    md = ModelData(models.square_heisenberg(), (4, 4), 1.0)
    md.struct.n_basis = 0
    with pytest.raises(_lib.CsmcError):
        _lib.plan(md)


@pytest.mark.parametrize("name,builder,shape,bc,ncol,structured", PLAN_CASES, ids=[c[0] for c in PLAN_CASES])
def test_library_tables_match_oracle_tables(name, builder, shape, bc, ncol, structured):
    """csmc_reference_tables (host-only closed form, what Lattice.bilinear_sites etc. expose) against the
    oracle's literal restatement of the reference constructor's findfirst scan."""
    md = ModelData(builder(), shape, 1.0, bc)
    lat = orc.OracleLattice(md, literal=True)
    for a, b in zip(_lib.reference_tables(md), lat.tables()):
        assert np.array_equal(a, b)


def test_jit_source_generation_and_nvrtc_compile():
    """The per-model kernel source is generated and compiled for sm_100a without a GPU (csmc_jit_check)."""
    md = ModelData(models.kitaev_honeycomb(), (8, 8), 1.0)
    src, log = _lib.jit_check(md, compile=True)
    assert 'extern "C" __global__' in src and "csmc_sweep_c0_u0" in src and "csmc_energy_c1" in src
    assert "struct Seg0" in src and "struct Seg1" in src
    # literal coefficients: K + J = -1 appears as an exact hexadecimal literal
    assert "(-0x1p+0)" in src
    md2 = ModelData(models.square_heisenberg(), (11, 13), 1.0)     # no periodic pattern -> no specialisation
    with pytest.raises(_lib.CsmcError, match="explicit-table"):
        _lib.jit_check(md2, compile=False)


def _check_skew_plan(n_rows, n_passes, reach, plan):
    """Replays a launch plan on per-row pass counters: every row runs every pass exactly once and in order, and
    whenever a row runs pass p each row within `reach` (periodic) has finished pass p - 1 and not finished pass
    p + 1 -- the neighbours a site reads (always of another colour) then hold exactly the values of the
    pass-by-pass order."""
    done = np.zeros(n_rows, np.int64)          # passes finished per row
    for p, row0, nrows in plan:
        assert 0 <= row0 and nrows >= 1 and row0 + nrows <= n_rows, (p, row0, nrows)
        rows = np.arange(row0, row0 + nrows)
        assert (done[rows] == p).all(), f"pass {p} out of order on rows {row0}..{row0 + nrows}"
        for k in range(1, reach + 1):
            for nb in ((rows - k) % n_rows, (rows + k) % n_rows):
                assert ((done[nb] == p) | (done[nb] == p + 1)).all(), f"pass {p}, rows {row0}+{nrows}: neighbour state"
        done[rows] = p + 1
    assert (done == n_passes).all()


@pytest.mark.parametrize("n_rows,n_passes,reach,budget", [(512, 22, 1, 85), (512, 44, 1, 120), (64, 4, 1, 16), (1024, 22, 2, 170),
                                                            (100, 6, 3, 40), (37, 2, 1, 5), (512, 22, 1, 43), (512, 22, 1, 44), (128, 12, 1, 85),
                                                            (32, 6, 1, 21), (64, 6, 1, 200), (40, 22, 1, 30)])
def test_time_skewed_strip_schedule_preserves_every_dependency(n_rows, n_passes, reach, budget):
    plan = _lib.skew_schedule(n_rows, n_passes, reach, budget)
    shift = (n_passes - 1) * reach
    if min(budget, (n_rows + 2 * shift) // 2) - 2 * shift < 1:
        assert len(plan) == 0          # the strips would vanish before the last pass: pass by pass
        return
    assert len(plan) > 0
    _check_skew_plan(n_rows, n_passes, reach, [tuple(int(v) for v in row) for row in plan])
    # the L2 working set: no launch is wider than the budget
    assert plan[:, 2].max() <= budget


def test_time_skewed_schedule_checker_rejects_a_wrong_plan():
    plan = [tuple(int(v) for v in row) for row in _lib.skew_schedule(64, 4, 1, 16)]
    unskewed = [(p, r0, n) for s in range(4) for p in range(4) for (r0, n) in [(16 * s, 16)]]   # strips that do not move
    with pytest.raises(AssertionError):
        _check_skew_plan(64, 4, 1, unskewed)
    with pytest.raises(AssertionError):
        _check_skew_plan(64, 4, 1, plan[:-1])


def test_skew_kernels_compile_and_default_source_is_unchanged(monkeypatch):
    md = ModelData(models.square_heisenberg(), (256, 256), 1.0)
    monkeypatch.delenv("CSMC_SKEW", raising=False)
    src, _ = _lib.jit_check(md, compile=False)
    assert "#define CSMC_SKEW" not in src.split("typedef unsigned int uint32_t;")[0]


def _site_rows(md, n_tile_rows):
    """CTA-tile row (along lattice dimension 0) of every site, reference site order (basis slowest, last dim fastest)."""
    shape = md.shape
    per_basis = int(np.prod(shape))
    idx = np.arange(md.n_sites) % per_basis
    i0 = idx // int(np.prod(shape[1:]))
    return i0 // (shape[0] // n_tile_rows)


@pytest.mark.parametrize("name,builder,shape,n_sweeps,budget", [
    ("square", models.square_heisenberg, (128, 64), 2, 9),
    ("honeycomb-J3", lambda: models.kitaev_honeycomb(J3=0.25), (64, 32), 3, 12),
    ("triangular-multispin", models.triangular_multispin, (256, 64), 2, 20),
])
def test_time_skewed_order_reproduces_colour_order_in_the_oracle(name, builder, shape, n_sweeps, budget):
    """Arithmetic-level check on the CPU: the oracle's overrelaxation run in the order of the time-skewed launch plan
    (strip by strip, the library's colouring and tile-row geometry) gives bit for bit the spins of the same sweeps run
    colour pass by colour pass; stationary strips (no skew) do not."""
    md = ModelData(builder(), shape, 1.0)
    usable, rows, reach, _ = _lib.skew_geometry(md)
    assert usable and rows >= 8
    colour, n_col, structured, _ = _lib.plan(md)
    assert structured
    lat = orc.OracleLattice(md)
    row_of = _site_rows(md, rows)
    sites = np.arange(1, md.n_sites + 1)

    s_ref = lat.randomize(seed=11)
    s_skew, s_bad = s_ref.copy(), s_ref.copy()
    colour_order = np.concatenate([sites[colour == c] for c in range(n_col)])
    lat.overrelax(s_ref, colour_order, n_sweeps)

    plan = _lib.skew_schedule(rows, n_sweeps * n_col, reach, budget)
    assert len(plan) > 0
    for p, row0, nrows in plan:
        sel = sites[(colour == p % n_col) & (row_of >= row0) & (row_of < row0 + nrows)]
        lat.overrelax(s_skew, sel, 1)
    assert np.array_equal(s_skew, s_ref)

    half = rows // 2                                   # two strips that do not move: dependencies are violated
    for r0, r1 in ((0, half), (half, rows)):
        for p in range(n_sweeps * n_col):
            sel = sites[(colour == p % n_col) & (row_of >= r0) & (row_of < r1)]
            lat.overrelax(s_bad, sel, 1)
    assert not np.array_equal(s_bad, s_ref)


def test_time_skewed_order_with_a_metropolis_sweep_in_the_oracle():
    """Same check for a cycle of 2 overrelaxation sweeps + 1 Metropolis sweep (counter-based Philox stream per site, so
    the visiting order cannot change the random numbers): spins and the accepted count are identical."""
    md = ModelData(models.kitaev_honeycomb(J3=0.25), (64, 32), 1.0)
    _, rows, reach, _ = _lib.skew_geometry(md)
    colour, n_col, _, _ = _lib.plan(md)
    lat = orc.OracleLattice(md)
    row_of = _site_rows(md, rows)
    sites = np.arange(1, md.n_sites + 1)
    kinds = ["or", "or", "metro"]
    T, seed = 0.6, 4242

    def run(spins, sel, kind):
        if kind == "or":
            lat.overrelax(spins, sel, 1)
            return 0
        return lat.metropolis_philox(spins, sel, T, seed, 0, 7)

    s_ref = lat.randomize(seed=3)
    s_skew = s_ref.copy()
    acc_ref = sum(run(s_ref, sites[colour == c], kinds[k]) for k in range(3) for c in range(n_col))
    acc_skew = 0
    plan = _lib.skew_schedule(rows, 3 * n_col, reach, 12)
    assert len(plan) > 0
    for p, row0, nrows in plan:
        sel = sites[(colour == p % n_col) & (row_of >= row0) & (row_of < row0 + nrows)]
        acc_skew += run(s_skew, sel, kinds[p // n_col])
    assert np.array_equal(s_skew, s_ref) and acc_skew == acc_ref and 0 < acc_ref < md.n_sites


@pytest.mark.parametrize("workload,replicas,colours", [("C2", 1, 2), ("C3", 8, 2), ("C3", 64, 2), ("C4", 16, 4), ("C5", 1, 4)])
def test_persistent_kernel_tiling_plan(workload, replicas, colours):
    """csmc_persist_check (host only): the tiling of the tile-resident persistent kernel for the BASELINE workloads on a
    B200 (148 SMs, 227 KiB of shared memory per CTA): the tiles cover the supercell grid exactly, one launch never needs
    more CTAs than there are SMs (co-residency of the cooperative launch), and the padded tiles of every class fit the
    shared memory."""
    from classicalspinmc.jl_b200 import workloads
    md, _ = workloads.workload_model(workload)
    info, src, _ = _lib.persist_check(md, replicas, compile=False)
    assert info["usable"]
    assert info["tiles"] == info["g0"] * info["g1"] and info["tiles"] * info["replicas_per_launch"] <= 148
    assert 1 <= info["replicas_per_launch"] <= replicas
    assert info["smem"] + 1024 <= 227 * 1024
    col, ncol, structured, _ = _lib.plan(md)
    assert ncol == colours and structured
    # supercell extents along the two tiled dimensions: lattice extent / colouring period; g tiles of extent w cover them
    m = re.search(r"tiles of (\d+) x (\d+) x (\d+) supercells per replica \(padded (\d+) x (\d+)\)", src)
    assert m and (int(m.group(1)), int(m.group(2))) == (info["w0"], info["w1"])
    per = [p for p in (1, 2) if md.shape[0] % p == 0]
    M0 = [md.shape[0] // p for p in per]
    assert any((info["g0"] - 1) * info["w0"] < M <= info["g0"] * info["w0"] for M in M0)
    assert "csmc_persist" in src and "st_release_gpu" in src and "persist_wait" in src


def test_persistent_kernel_compiles_and_rejects_what_it_cannot_tile():
    from classicalspinmc.jl_b200 import workloads
    info, _, log = _lib.persist_check(ModelData(models.kitaev_honeycomb(J3=0.25), (64, 48), 1.0), 3, compile=True)
    assert info["usable"] and "error" not in log.lower()
    # a lattice whose replica does not fit the SMs' shared memory: 384 MiB of spins
    md, _ = workloads.workload_model("C2", 4096)
    assert not _lib.persist_check(md, 1, compile=False)[0]["usable"]
    # open boundaries are left to the pass kernels
    assert not _lib.persist_check(ModelData(models.square_heisenberg(), (64, 64), 1.0, "open"), 1, compile=False)[0]["usable"]
    # a device with few SMs still gets a plan (more, smaller launches are not needed: tiles grow until they fit)
    few = _lib.persist_check(ModelData(models.square_heisenberg(), (256, 256), 1.0), 4, n_sms=16, compile=False)[0]
    assert few["usable"] and few["tiles"] * few["replicas_per_launch"] <= 16
