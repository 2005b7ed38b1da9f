"""Pins the CPU oracle against every golden value / known answer the reference's own tests hold
for the hot path:
  test/latticetests.jl:6   |s| == S for random initial spins
  test/latticetests.jl:17  Zeeman: total_energy == -1.0
  test/latticetests.jl:18  get_local_field == (-1, -0, -0)
  test/latticetests.jl:30  bilinear: total_energy / size == -2.0
  test/mctests.jl:49,57    annealed Kitaev-Gamma honeycomb: round(E/N, 4) == -0.6444
and checks the closed-form table construction against a literal restatement of the reference's
`findfirst` scan (src/lattice.jl:196,228-229,273-275) on small lattices.
"""
import numpy as np
import pytest

import classicalspinmc.jl_b200 as csm
from classicalspinmc.jl_b200._abi import ModelData
from oracle import oracle as orc
from tests import models


def test_norm_spin():
    # test/latticetests.jl:3-7
    md = ModelData(_with_basis(csm.Square()), (2, 2), 1.0)
    lat = orc.OracleLattice(md)
    s = lat.randomize(seed=42)
    assert np.all(np.round(np.linalg.norm(s, axis=1), 9) == 1.0)


def _with_basis(uc):
    if len(uc.basis) == 0:
        csm.addBasisSite(uc, np.zeros(uc.D))  # src/lattice.jl:68-70
    return uc


def test_energy_zeeman_exact():
    # test/latticetests.jl:10-19
    uc = csm.Square()
    h = np.array([1.0, 0.0, 0.0])
    csm.addZeemanCoupling(uc, 1, h)
    lat = orc.OracleLattice(ModelData(_with_basis(uc), (1, 1), 1.0))
    spins = np.array([[1.0, 0.0, 0.0]])
    assert lat.total_energy(spins) == -1.0
    f = lat.local_field(spins, 1)
    assert tuple(f) == (-1.0, -0.0, -0.0)


def test_energy_bilinear_exact():
    # test/latticetests.jl:21-31
    uc = models.square_heisenberg(J=-1.0, h=None)
    lat = orc.OracleLattice(ModelData(_with_basis(uc), (2, 2), 1.0))
    spins = np.tile(np.array([1.0, 0.0, 0.0]), (4, 1))
    assert lat.total_energy(spins) / lat.N == -2.0
    bil, _, _ = lat.tables()
    # on L=2 the +-x neighbours are the same site counted twice
    assert bil.tolist()[0] == [3, 3, 2, 2]


@pytest.mark.parametrize("alg", [0, 1])
def test_annealing_ground_state(alg):
    # test/mctests.jl:1-58: honeycomb L=4, K=-1, G=0.2, Gp=-0.02, h=0.1*[111]/sqrt3
    uc = models.kitaev_honeycomb()
    lat = orc.OracleLattice(ModelData(uc, (4, 4), 1.0))
    spins = lat.randomize(seed=7 + alg)
    T0, T = 1.0, 1e-7
    temps = orc.annealing_temperatures(T, lambda x: T0 * 0.9 ** x, T0)
    assert len(temps) == 153
    lat.simulated_annealing(spins, temps, int(1e4), 10, alg=alg, seed=11 + alg)
    if alg == 0:
        lat.deterministic_updates(spins, int(1e6), seed=5)
    E = lat.total_energy(spins) / lat.N
    assert round(E, 4) == -0.6444


CASES = [
    ("square", lambda: models.square_heisenberg(), (3, 4), "periodic"),
    ("square-open", lambda: models.square_heisenberg(), (3, 4), "open"),
    ("honeycomb", lambda: models.kitaev_honeycomb(J3=0.3), (3, 5), "periodic"),
    ("pyrochlore", lambda: models.pyrochlore_local(), (2, 3, 2), "periodic"),
    ("triangular-multispin", lambda: models.triangular_multispin(), (4, 4), "periodic"),
    ("triangular-multispin-open", lambda: models.triangular_multispin(), (3, 4), "open"),
    ("mixed-basis", lambda: models.mixed_basis_multispin(), (3, 4), "periodic"),
    ("mixed-basis-open", lambda: models.mixed_basis_multispin(), (3, 3), "open"),
]


@pytest.mark.parametrize("name,builder,shape,bc", CASES, ids=[c[0] for c in CASES])
def test_closed_form_tables_match_literal_findfirst(name, builder, shape, bc):
    md = ModelData(builder(), shape, 1.0, bc)
    fast = orc.OracleLattice(md, literal=False)
    slow = orc.OracleLattice(md, literal=True)
    for a, b in zip(fast.tables(), slow.tables()):
        assert np.array_equal(a, b)
    assert np.array_equal(fast.bilinear_matrices(), slow.bilinear_matrices())


@pytest.mark.parametrize("name,builder,shape,bc", CASES, ids=[c[0] for c in CASES])
def test_site_energy_field_consistency(name, builder, shape, bc):
    """Self-consistency of the (unpinned) multi-spin restatement: with every perspective of a term
    registered, e_i is linear in s_i through the field: e_i = s_i.(F_i + h_i - O s_i) - s_i.h_i
    (src/hamiltonian.jl:3-67 vs :139-196), and total_energy equals the weighted site sums."""
    md = ModelData(builder(), shape, 0.7, bc)
    lat = orc.OracleLattice(md)
    s = lat.randomize(seed=3)
    F = lat.local_field_all(s)
    e = lat.site_energy_all(s)
    from classicalspinmc.jl_b200._abi import resolve_field_onsite
    nb = md.n_basis
    cells = lat.N // nb
    basis_of = np.repeat(np.arange(nb), cells)
    h = md.field[basis_of]
    O = md.onsite[basis_of].reshape(-1, 3, 3)
    Os = np.einsum("nab,nb->na", O, s)
    e_from_field = np.einsum("na,na->n", s, F - Os)
    assert np.allclose(e, e_from_field, rtol=0, atol=1e-12)


def test_multispin_perspectives_sum_rule():
    """Sum_i e3_i / 3 and Sum_i e4_i / 4 equal the direct per-cell sums."""
    L = (4, 4)
    uc = models.triangular_multispin(J=0.0)
    # J=0 drops the bilinear terms entirely (src/unit_cell.jl:49)
    assert len(uc.bilinear) == 0
    md = ModelData(uc, L, 1.0)
    lat = orc.OracleLattice(md)
    s = lat.randomize(seed=9)
    C3 = np.random.default_rng(7).uniform(-0.1, 0.1, (3, 3, 3))
    R4 = np.random.default_rng(8).uniform(-0.05, 0.05, (3, 3, 3, 3))
    S = s.reshape(L[0], L[1], 3)
    direct = 0.0
    for x in range(L[0]):
        for y in range(L[1]):
            s0 = S[x, y]; s1 = S[(x + 1) % L[0], y]; s2 = S[x, (y + 1) % L[1]]
            s3 = S[(x + 1) % L[0], (y + 1) % L[1]]
            direct += np.einsum("abc,a,b,c->", C3, s0, s1, s2)
            direct += np.einsum("abcd,a,b,c,d->", R4, s0, s1, s2, s3)
    assert abs(lat.total_energy(s) - direct) < 1e-12


# ---- committed fixtures (tests/golden/, generated by tests/golden/make_vectors.py) -----------------------------
def _golden():
    import importlib.util
    import os
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    spec = importlib.util.spec_from_file_location("make_vectors", os.path.join(here, "make_vectors.py"))
    mv = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mv)
    return mv, np.load(os.path.join(here, "oracle_vectors.npz")), os.path.join(here, "reference_known_answers.json")


def test_known_answers_file_lists_what_this_module_pins():
    import json
    _, _, path = _golden()
    answers = json.load(open(path))["answers"]
    cites = " ".join(a["cite"] for a in answers)
    for needle in ("latticetests.jl:3-7", "latticetests.jl:10-17", "latticetests.jl:18", "latticetests.jl:21-30", "mctests.jl:36-49", "mctests.jl:52-58"):
        assert needle in cites
    assert [a["value"] for a in answers if "mctests" in a["cite"]] == [-0.6444, -0.6444]


def test_oracle_reproduces_the_committed_vectors():
    """The fixtures the GPU parity tests read were produced by this oracle: regenerating them gives the same
    numbers (<= 1e-13: the library may be rebuilt by another compiler version)."""
    mv, z, _ = _golden()
    assert {k.split("/")[0] for k in z.files} == set(mv.CASES)
    for name in mv.CASES:
        fresh = mv.build(name)
        for k, v in fresh.items():
            ref = z[f"{name}/{k}"]
            assert ref.shape == np.asarray(v).shape, (name, k)
            assert np.abs(ref - v).max() <= 1e-13 * max(1.0, np.abs(ref).max()), (name, k)


@pytest.mark.parametrize("name,builder,shape,S,frac", [
    ("square-96", lambda: models.square_heisenberg(), (96, 96), 1.0, 0.97),
    ("honeycomb-J3-48", lambda: models.kitaev_honeycomb(J3=0.25), (48, 48), 1.0, 0.9),
    ("pyrochlore-8", lambda: models.pyrochlore_local(), (8, 8, 8), 0.5, 0.5),
    ("triangular-multispin-32", lambda: models.triangular_multispin(), (32, 32), 1.0, 0.3),
])
def test_forward_error_bound_covers_a_differently_rounded_build(name, builder, shape, S, frac):
    """The per-site bound the full-size GPU parity tests assert (orc_sweep_tracked: |difference| <= 1e-12 * S * kappa_i
    on every site) must cover two implementations of the same updates that differ only in rounding.  Here: the
    bit-exactness build of the oracle (-ffp-contract=off) against its performance build (-O3 -march=native, FMA
    contraction) in the library's colour order, sweep by sweep from identical inputs and over two sweeps in one go; and
    the bound must stay informative (kappa <= 1, i.e. the plain 1e-12, on most sites of a sweep)."""
    from classicalspinmc.jl_b200 import _lib
    md = ModelData(builder(), shape, S)
    a, b = orc.OracleLattice(md), orc.OracleLattice(md, fast=True)
    order = (np.argsort(_lib.plan(md)[0], kind="stable") + 1).astype(np.int64)
    s = a.randomize(seed=77)
    s0 = s.copy()
    for kind in (0, 0, 1):
        t = s.copy()
        kappa = np.zeros(a.N)
        a.sweep_tracked(s, order, kind, kappa)
        (b.overrelax if kind == 0 else b.deterministic)(t, order, 1)
        assert np.all(np.abs(s - t).max(axis=1) <= 1e-12 * S * kappa)
        assert (kappa <= 1.0).mean() > frac
    s, t, kappa = s0.copy(), s0.copy(), np.zeros(a.N)
    for _ in range(2):
        a.sweep_tracked(s, order, 0, kappa)
    b.overrelax(t, order, 2)
    assert np.all(np.abs(s - t).max(axis=1) <= 1e-12 * S * kappa)


def _numpy_multispin_reference(uc, shape, bc, spins):
    """An independent, literal numpy restatement of the reference's cubic / quartic code — table construction with
    `permutedims` (src/lattice.jl:209-283, first matching branch) and the `@einsum` lines of get_local_field and energy
    (src/hamiltonian.jl:46-48, 62-64, 180, 193) — for 2-D lattices: the arithmetic that lives in Einsum.jl, for which
    the reference holds no test vector.  Returns (cubic + quartic part of the field [N, 3], of the site energy [N])."""
    import itertools
    nb = len(uc.basis)
    L1, L2 = shape
    indices = sorted((b, i1, i2) for b in range(1, nb + 1) for i1 in range(1, L1 + 1) for i2 in range(1, L2 + 1))   # :29-33
    lookup = {t: p for p, t in enumerate(indices)}

    def BC(index, off):                                                      # :101-109
        out = []
        for d, L in enumerate(shape):
            v = index[1 + d] + off[d]
            if bc == "periodic":
                v = (v - 1) % L + 1
            out.append(v)
        return tuple(out)

    def find(b, cell):
        return lookup.get((b,) + cell)                                       # findfirst(...) or nothing (open bc)

    N = len(indices)
    F = np.zeros((N, 3))
    E = np.zeros(N)
    for p, index in enumerate(indices):
        s = spins[p]
        for (b1, b2, b3, J, oj, ok) in uc.cubic:                            # :209-237
            J = np.asarray(J)
            oj, ok = np.array(oj), np.array(ok)
            if index[0] not in (b1, b2, b3):
                continue
            if b1 == index[0]:
                bj, bk = b2, b3
            elif b2 == index[0]:
                bj, bk = b1, b3
                ok = ok - oj
                oj = oj * -1
                J = np.transpose(J, (1, 0, 2))                               # permutedims(J, [2, 1, 3])
            else:
                bj, bk = b2, b1
                oj = oj - ok
                ok = ok * -1
                J = np.transpose(J, (2, 1, 0))                               # permutedims(J, [3, 2, 1])
            j, k = find(bj, BC(index, oj)), find(bk, BC(index, ok))
            if j is None or k is None:
                continue
            sj, sk = spins[j], spins[k]
            for x in range(3):
                F[p, x] += np.einsum("ab,a,b->", J[x], sj, sk)               # @einsum Hx += C[1, a, b] * sj[a] * sk[b]
            E[p] += np.einsum("abc,a,b,c->", J, s, sj, sk)                   # @einsum E += C[a, b, c] * s[a] * sj[b] * sk[c]
        for (b1, b2, b3, b4, J, oj, ok, ol) in uc.quartic:                   # :243-283
            J = np.asarray(J)
            oj, ok, ol = np.array(oj), np.array(ok), np.array(ol)
            if index[0] not in (b1, b2, b3, b4):
                continue
            if b1 == index[0]:
                bj, bk, bl = b2, b3, b4
            elif b2 == index[0]:
                bj, bk, bl = b1, b3, b4
                oj = oj * -1
                ok = ok + oj
                ol = ol + oj
                J = np.transpose(J, (1, 0, 2, 3))                            # [2, 1, 3, 4]
            elif b3 == index[0]:
                bj, bk, bl = b2, b1, b4
                ok = ok * -1
                ol = ol + ok
                oj = oj + ok
                J = np.transpose(J, (2, 1, 0, 3))                            # [3, 2, 1, 4]
            else:
                bj, bk, bl = b2, b3, b1
                ol = ol * -1
                oj = oj + ol
                ok = ok + ol
                J = np.transpose(J, (3, 1, 2, 0))                            # [4, 2, 3, 1]
            j, k, l = find(bj, BC(index, oj)), find(bk, BC(index, ok)), find(bl, BC(index, ol))
            if j is None or k is None or l is None:
                continue
            sj, sk, sl = spins[j], spins[k], spins[l]
            for x in range(3):
                F[p, x] += np.einsum("abc,a,b,c->", J[x], sj, sk, sl)        # @einsum Hx += R[1, a, b, c] * sj[a] * sk[b] * sl[c]
            E[p] += np.einsum("abcd,a,b,c,d->", J, s, sj, sk, sl)
    return F, E


@pytest.mark.parametrize("name,builder,shape,bc,S", [
    ("mixed-basis-4x6", lambda: models.mixed_basis_multispin(), (4, 6), "periodic", 0.8),
    ("mixed-basis-open-5x3", lambda: models.mixed_basis_multispin(), (5, 3), "open", 0.8),
    ("triangular-multispin-6x4", lambda: models.triangular_multispin(), (6, 4), "periodic", 1.0),
])
def test_cubic_and_quartic_terms_against_a_literal_numpy_restatement(name, builder, shape, bc, S):
    """Pins the oracle's cubic / quartic arithmetic (every perspective branch, both boundary conditions) against an
    independent numpy restatement of the reference's lines; the bilinear / on-site / Zeeman part is taken out by
    evaluating the oracle on the same model without multi-spin terms."""
    uc = builder()
    md = ModelData(uc, shape, S, bc)
    lat = orc.OracleLattice(md)
    s = lat.randomize(seed=9)
    uc0 = builder()
    uc0.cubic.clear()
    uc0.quartic.clear()
    lat0 = orc.OracleLattice(ModelData(uc0, shape, S, bc))
    F = lat.local_field_all(s) - lat0.local_field_all(s)
    E = lat.site_energy_all(s) - lat0.site_energy_all(s)
    F_ref, E_ref = _numpy_multispin_reference(uc, shape, bc, s)
    assert np.abs(F_ref).max() > 1e-3 and np.abs(E_ref).max() > 1e-3          # the terms are really there
    assert np.abs(F - F_ref).max() <= 1e-13 * max(1.0, np.abs(F_ref).max())
    assert np.abs(E - E_ref).max() <= 1e-13 * max(1.0, np.abs(E_ref).max())
