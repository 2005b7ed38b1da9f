"""CPU tests of the host-side mirror of the reference's Julia layer: UnitCell / Lattice tables,
parameter buffer, observables (binning), output-file layout, parallel-tempering bookkeeping."""
import math
import os
import warnings

import numpy as np
import pytest

import classicalspinmc.jl_b200 as csm
from classicalspinmc.jl_b200 import hdf5 as h5
from classicalspinmc.jl_b200 import parallel
from classicalspinmc.jl_b200._abi import ModelData, resolve_field_onsite
from classicalspinmc.jl_b200.observables import ErrorPropagator, _specific_heat
from oracle import oracle as orc
from tests import models


def test_zero_couplings_are_dropped():
    # src/unit_cell.jl:38,49,60,72
    uc = csm.Square()
    csm.addBilinear(uc, 1, 1, np.zeros((3, 3)), (1, 0))
    csm.addOnSite(uc, 1, np.zeros((3, 3)))
    csm.addCubic(uc, 1, 1, 1, np.zeros((3, 3, 3)))
    csm.addQuartic(uc, 1, 1, 1, 1, np.zeros((3, 3, 3, 3)))
    assert not uc.bilinear and not uc.onsite and not uc.cubic and not uc.quartic


def test_lattice_adds_default_basis_and_normalises_spins():
    # src/lattice.jl:68-70 and test/latticetests.jl:3-7
    uc = csm.Square()
    lat = csm.Lattice((2, 2), uc, 1.0)
    assert len(uc.basis) == 1 and lat.size == 4 and lat.spins.shape == (3, 4)
    assert np.all(np.round(np.linalg.norm(lat.spins, axis=0), 9) == 1.0)
    fm = csm.Lattice((3, 3), csm.Square(), 0.5, initialCondition="fm")
    assert np.allclose(fm.spins, fm.spins[:, :1]) and np.allclose(np.linalg.norm(fm.spins, axis=0), 0.5)
    with pytest.raises(ValueError, match="Invalid boundary condition option"):
        csm.Lattice((2, 2), csm.Square(), 1.0, bc="twisted")


def test_site_order_and_positions():
    # src/lattice.jl:29-51: basis slowest, last lattice index fastest
    uc = csm.Honeycomb()
    lat = csm.Lattice((2, 3), uc, 1.0)
    idx = csm.lattice.site_indices((2, 3), 2)
    assert idx[0] == (1, 1, 1) and idx[1] == (1, 1, 2) and idx[3] == (1, 2, 1) and idx[6] == (2, 1, 1)
    a1, a2 = uc.lattice_vectors
    for p, (b, i, j) in enumerate(idx):
        expect = (i - 1) * a1 + (j - 1) * a2 + uc.basis[b - 1]
        assert np.allclose(lat.site_positions[:, p], expect)


@pytest.mark.parametrize("bc", ["periodic", "open"])
def test_lattice_tables_match_oracle(bc):
    """Public table fields (lat.bilinear_sites, ..., src/lattice.jl:14-22) vs the oracle's literal
    restatement of the reference constructor."""
    uc = models.mixed_basis_multispin()
    lat = csm.Lattice((3, 4), uc, 0.8, bc=bc)
    o = orc.OracleLattice(ModelData(uc, (3, 4), 0.8, bc), literal=True)
    bil, cub, quar = o.tables()
    assert np.array_equal(lat.bilinear_sites, bil)
    assert np.array_equal(lat.cubic_sites, cub)
    assert np.array_equal(lat.quartic_sites, quar)
    assert np.array_equal(lat.bilinear_matrices.reshape(lat.size, -1, 9), o.bilinear_matrices())
    # field / onsite per site
    md = lat._model
    b_of = np.repeat(np.arange(md.n_basis), lat.size // md.n_basis)
    assert np.array_equal(lat.field, md.field[b_of])
    # energy from the public tables (plain numpy, bilinear part) equals the oracle's bilinear energy
    s = np.ascontiguousarray(lat.spins.T)
    e2 = 0.0
    for p in range(lat.size):
        for n in range(bil.shape[1]):
            j = lat.bilinear_sites[p, n]
            if j:
                e2 += s[p] @ lat.bilinear_matrices[p, n] @ s[j - 1]
    uc2 = csm.Honeycomb()
    for t in uc.bilinear:
        csm.addBilinear(uc2, *t)
    o2 = orc.OracleLattice(ModelData(uc2, (3, 4), 0.8, bc))
    assert abs(o2.total_energy(s) - e2 / 2) < 1e-12


def test_field_resolution_quirk():
    # src/lattice.jl:128-133 indexes the term list by basis number
    uc = csm.Honeycomb()
    csm.addZeemanCoupling(uc, 2, np.array([0.0, 0.0, 2.0]))
    csm.addZeemanCoupling(uc, 1, np.array([1.0, 0.0, 0.0]))
    f, _ = resolve_field_onsite(uc)
    assert np.array_equal(f, [[1.0, 0, 0], [0, 0, 2.0]])
    uc = csm.Honeycomb()
    csm.addZeemanCoupling(uc, 2, np.array([0.0, 0.0, 2.0]))   # only basis 2: BoundsError in Julia
    with pytest.raises(IndexError):
        resolve_field_onsite(uc)


def test_params_buffer_defaults_and_warning():
    d = {"t_thermalization": 100, "overrelaxation": 10}
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        p = csm.MCParamsBuffer(d)
    assert any("not a valid MC parameter" in str(x.message) for x in w)     # README.md:49 key is ignored
    assert p == (100, 1, 1, 1, 1, 10, 0, 0)
    assert d["swap_rate"] == 1                                                # defaults are written back


def test_error_propagator_matches_direct_statistics():
    rng = np.random.default_rng(0)
    x = rng.normal(3.0, 2.0, 4096)
    ep = ErrorPropagator(2)
    for v in x:
        ep.push(v, v * v)
    assert len(ep) == 4096
    assert abs(ep.mean(1) - x.mean()) < 1e-12
    assert abs(ep.mean(2) - (x * x).mean()) < 1e-9
    # uncorrelated data: the binned standard error agrees with the naive one
    assert abs(ep.std_error(1) / (x.std(ddof=1) / np.sqrt(len(x))) - 1) < 0.25
    lvl = ep.reliable_level()
    assert ep.count[lvl] >= 32 and (lvl + 1 >= len(ep.count) or ep.count[lvl + 1] < 32)
    # level l holds means of 2^l consecutive samples
    assert ep.count[3] == 4096 // 8
    assert abs(ep.sums1D[3, 0] / ep.count[3] - x.mean()) < 1e-12
    # correlated series: binning inflates the error bar
    y = np.repeat(rng.normal(0, 1, 512), 8)
    ep2 = ErrorPropagator(2)
    for v in y:
        ep2.push(v, v * v)
    assert ep2.std_error(1) > 1.8 * (y.std(ddof=1) / np.sqrt(len(y)))
    c, dc = _specific_heat(ep, 0.5, 100)
    assert abs(c - (np.mean(x * x) - x.mean() ** 2) / 0.25 / 100) < 1e-9 and dc > 0


def test_output_files_roundtrip(tmp_path):
    """test/h5tests.jl:5-46 restated for this host layer (plus the key strings of src/hdf5.jl:60-74,
    which the reference's own test does not exercise)."""
    uc = models.mixed_basis_multispin()
    lat = csm.Lattice((2, 3), uc, 0.8)
    out = str(tmp_path) + "/"
    mc = csm.MonteCarlo(0.3, lat, {"t_thermalization": 10, "overrelaxation_rate": 5}, outpath=out,
                        inparams={"K": -1.0, "tag": 7})
    assert os.path.isfile(out + "configuration.h5.params") and os.path.isfile(out + "configuration_0.h5")
    f = h5._open(out + "configuration.h5.params", "r")
    keys = h5._keys(f, "unit_cell/bilinear")
    assert "(1,2),(0, -1)" in keys and "(2,1),(1, 0)" in keys
    assert h5._get_attr(f, "t_thermalization") == 10 and h5._get_attr(f, "K") == -1.0
    lat2 = csm.read_lattice(f)
    f.close()
    assert lat2.shape == lat.shape and lat2.S == lat.S and lat2.bc == lat.bc
    u2 = lat2.unit_cell
    assert np.allclose(np.stack(u2.lattice_vectors), np.stack(uc.lattice_vectors))
    assert len(u2.bilinear) == len(uc.bilinear) and len(u2.cubic) == 2 and len(u2.quartic) == 1
    key = lambda t: (t[0], t[1], tuple(t[3]))
    for a, b in zip(sorted(u2.bilinear, key=key), sorted(uc.bilinear, key=key)):
        assert a[:2] == b[:2] and tuple(a[3]) == tuple(b[3]) and np.array_equal(a[2], b[2])
    assert np.array_equal(sorted(u2.quartic)[0][4], uc.quartic[0][4])
    # term order follows the (alphabetical) key order after a round trip, as with HDF5 itself
    assert np.array_equal(np.sort(lat2.bilinear_sites, axis=1), np.sort(lat.bilinear_sites, axis=1))
    # configuration file: spins as (N, 3) (util/load.py:88-93), checkpoint overwrite, resume
    mc.lattice.spins[:] = 0.25
    csm.write_MC_checkpoint(mc)
    lat3 = csm.Lattice((2, 3), uc, 0.8)
    csm.read_spin_configuration(lat3, out + "configuration_0.h5")
    assert np.all(lat3.spins == 0.25)
    g = h5._open(out + "configuration_0.h5", "r")
    assert np.asarray(h5._get(g, "spins")).shape == (lat.size, 3)
    assert h5._get_attr(g, "T") == 0.3
    g.close()
    # MonteCarlo deep-copies the lattice (src/monte_carlo.jl:74; test/mctests.jl:45,54 rely on it)
    assert not np.all(lat.spins == 0.25)


def test_pairing_and_slot_bookkeeping():
    # src/monte_carlo.jl:311-317
    assert parallel.pairing(5, 0) == [(0, 1), (2, 3)]
    assert parallel.pairing(5, 1) == [(1, 2), (3, 4)]
    assert parallel.pairing(2, 1) == []
    s = parallel.apply_exchanges(np.arange(4), [0, 2])
    assert s.tolist() == [1, 0, 3, 2]
    s = parallel.apply_exchanges(s, [1])       # replicas now in slots 1,2 are 0 and 3
    assert s.tolist() == [2, 0, 3, 1]
    assert sorted(s.tolist()) == [0, 1, 2, 3]
    assert parallel.owner_of_replica(5, [4, 4]) == 1 and parallel.owner_of_replica(3, [4, 4]) == 0


def test_unsupported_variants_are_loud():
    uc = models.square_heisenberg()
    lat = csm.Lattice((2, 2), uc, 1.0)
    with pytest.raises(ValueError, match="No momentum vectors"):
        csm.MonteCarlo(1.0, lat, {}, corr=True)
    mc = csm.MonteCarlo(1.0, lat, {})
    with pytest.raises(NotImplementedError):
        csm.parallel_tempering(mc, alg=csm.MetropolisConstraintAdaptive())


def test_reciprocal_space_helpers():
    """src/reciprocal.jl: b_i . a_j = 2 pi delta_ij; commensurate wavevectors in Iterators.product order
    (first dimension fastest); k-paths made of allowed wavevectors between high-symmetry points."""
    for uc in (csm.Triangular(), csm.Honeycomb(), csm.Pyrochlore(), csm.FCC()):
        b = csm.reciprocal(*uc.lattice_vectors)
        assert np.allclose(np.stack(b) @ np.stack(uc.lattice_vectors).T, 2 * np.pi * np.eye(uc.D), atol=1e-12)
    uc = csm.Triangular()
    b1, b2 = csm.reciprocal(*uc.lattice_vectors)
    ks = csm.get_allowed_wavevectors(uc, (4, 6))
    assert ks.shape == (2, 5 * 7)
    assert np.allclose(ks[:, 1], b1 / 4) and np.allclose(ks[:, 5], b2 / 6) and np.allclose(ks[:, -1], b1 + b2)
    ks2 = csm.get_allowed_wavevectors(uc, (4, 6), min=-1, max=1)
    assert ks2.shape == (2, 9 * 13) and np.allclose(ks2[:, 0], -b1 - b2)
    # every allowed wavevector is a Bloch vector of the 4 x 6 torus: exp(i k . (L_d a_d)) == 1
    a1, a2 = uc.lattice_vectors
    assert np.allclose(np.exp(1j * (ks.T @ (4 * a1))), 1.0) and np.allclose(np.exp(1j * (ks.T @ (6 * a2))), 1.0)
    plane = csm.get_k_plane(uc, (4, 4), min=-1, max=1)
    assert plane.shape[0] == 2 and np.all(np.abs(plane) <= 2 * np.pi + 1e-6) and plane.shape[1] < 9 * 9
    hsp = {"G": np.zeros(2), "M": 0.5 * b1, "K": (2 * b1 + b2) / 3}
    count, kpath = csm.get_k_path(uc, hsp, ["G", "M", "K", "G"], (12, 12))
    assert count.tolist() == [0, 6, 8, 12] and kpath.shape == (2, 13)
    assert np.allclose(kpath[:, 0], 0) and np.allclose(kpath[:, 6], hsp["M"]) and np.allclose(kpath[:, 8], hsp["K"])
    assert np.allclose(kpath[:, 12], 0)
    line = csm.get_k_path(uc, np.array([1.0, 0.0]), (6, 6))
    assert line.shape[0] == 2 and line.shape[1] > 0 and np.allclose(line[1], 0.0, atol=1e-9)


def test_on_disk_orientation_is_what_hdf5_jl_produces(tmp_path):
    """HDF5.jl writes a column-major Julia array with its dataspace dimensions reversed, so a C-order reader sees the
    transpose of every array (src/hdf5.jl:39-40,60,67,74,229-236; util/load.py:88-93 relies on it for ``spins``).
    Checked on a non-symmetric cell: honeycomb lattice vectors, two basis sites, a DM-type (antisymmetric) bilinear
    matrix, a rank-3 tensor, the D x N site positions and a 9 x N_k structure factor."""
    from classicalspinmc.jl_b200 import hdf5 as h5
    uc = csm.Honeycomb()
    J = np.array([[1.0, 2.0, 3.0], [-2.0, 4.0, 5.0], [-3.0, -5.0, 6.0]])      # J[r, c] with J != J^T
    csm.addBilinear(uc, 1, 2, J, (0, -1))
    C = np.arange(27, dtype=float).reshape(3, 3, 3) + 1.0                      # C[a, b, c]
    csm.addCubic(uc, 1, 2, 1, C, (0, 0), (1, 0))
    lat = csm.Lattice((3, 2), uc, 1.0)
    mc = csm.MonteCarlo(1.0, lat, {"t_thermalization": 2}, outpath=str(tmp_path) + "/", outprefix="orient")
    f = h5._open(str(tmp_path) + "/orient.h5.params", "r")
    D, nb = 2, 2
    lv_disk = np.asarray(h5._get(f, "unit_cell/lattice_vectors"))
    # Julia: lattice_vectors[:, i] = a_i  ->  on disk row i = a_i
    assert lv_disk.shape == (D, D) and all(np.allclose(lv_disk[i], uc.lattice_vectors[i]) for i in range(D))
    basis_disk = np.asarray(h5._get(f, "unit_cell/basis"))
    assert basis_disk.shape == (D, nb)                                         # Julia n_basis x D  ->  disk D x n_basis
    assert all(np.allclose(basis_disk[:, b], uc.basis[b]) for b in range(nb))
    key = "unit_cell/bilinear/(1,2),(0, -1)"
    assert np.array_equal(np.asarray(h5._get(f, key)), J.T)                    # disk[c, r] = J[r, c]
    ckey = [k for k in h5._keys(f, "unit_cell/cubic")][0]
    assert np.array_equal(np.asarray(h5._get(f, "unit_cell/cubic/" + ckey)), C.transpose(2, 1, 0))
    f.close()
    g = h5._open(str(tmp_path) + "/orient_0.h5", "r")
    assert np.asarray(h5._get(g, "spins")).shape == (lat.size, 3)
    assert np.asarray(h5._get(g, "site_positions")).shape == (lat.size, D)
    g.close()
    # and the reader undoes it: every field of the unit cell and the lattice comes back
    lat2 = h5.read_lattice(str(tmp_path) + "/orient.h5.params")
    uc2 = lat2.unit_cell
    assert all(np.allclose(a, b) for a, b in zip(uc2.lattice_vectors, uc.lattice_vectors))
    assert len(uc2.basis) == nb and all(np.allclose(a, b) for a, b in zip(uc2.basis, uc.basis))
    assert np.array_equal(uc2.bilinear[0][2], J) and np.array_equal(uc2.cubic[0][3], C)
    assert np.allclose(lat2.site_positions, lat.site_positions)
    S = np.arange(9 * 4, dtype=float).reshape(9, 4)
    ks = np.arange(2 * 4, dtype=float).reshape(2, 4)
    g = h5._open(str(tmp_path) + "/orient_0.h5", "r+")
    h5.overwrite_keys(g, {"spin_correlations/SSF": S, "spin_correlations/SSF_momentum": ks})
    g.close()
    g = h5._open(str(tmp_path) + "/orient_0.h5", "r")
    assert np.asarray(h5._get(g, "spin_correlations/SSF")).shape == (4, 9)     # Julia 9 x N_k
    assert np.array_equal(h5._get_jl(g, "spin_correlations/SSF"), S)
    g.close()


def test_log_binning_hand_computed_known_answers():
    """The logarithmic binning behind specific_heat / susceptibility error bars (src/observables.jl:32-63 through
    BinningAnalysis.jl's ErrorPropagator) on a series small enough to do by hand: x = 1..8, pushed as (x, x^2).
      level 0: 8 values          mean 9/2, unbiased variance 6      -> std error sqrt(6/8)
      level 1: 3/2 7/2 11/2 15/2 (pair means)  variance 20/3        -> std error sqrt((20/3)/4)
      level 2: 5/2 13/2                         variance 8           -> std error sqrt(8/2) = 2
      level 3: 9/2 (a single bin: no variance)
    second argument x^2: level-0 mean 204/8 = 51/2; cov(x, x^2) at level 0 = (1296 - 36 * 204 / 8) / 7 = 54.
    First-order propagation of f = <x^2> - <x>^2 (the numerator of c and chi): gradient (-2 <x>, 1) = (-9, 1), so
    var f = 81 var(x) - 18 cov(x, x^2) + var(x^2) with var(x^2) = (8772 - 204^2 / 8) / 7 = 510: 81*6 - 18*54 + 510 = 24."""
    from classicalspinmc.jl_b200.observables import ErrorPropagator, std_error_tweak
    ep = ErrorPropagator(2)
    for x in range(1, 9):
        ep.push(float(x), float(x * x))
    assert ep.count[:5].tolist() == [8, 4, 2, 1, 0]
    assert ep.mean(1) == 4.5 and ep.mean(2) == 25.5
    assert np.isclose(ep.covmat(0)[0, 0], 6.0) and np.isclose(ep.covmat(1)[0, 0], 20.0 / 3.0) and np.isclose(ep.covmat(2)[0, 0], 8.0)
    assert np.isclose(ep.std_error(1, 0), math.sqrt(6.0 / 8.0))
    assert np.isclose(ep.std_error(1, 1), math.sqrt(20.0 / 3.0 / 4.0))
    assert np.isclose(ep.std_error(1, 2), 2.0)
    assert math.isnan(ep.std_error(1, 3))
    assert np.isclose(ep.covmat(0)[0, 1], 54.0) and np.isclose(ep.covmat(0)[1, 1], 510.0)
    grad = lambda m: np.array([-2.0 * m[0], 1.0])
    assert np.isclose(ep.var(grad, 0), 24.0)
    assert np.isclose(std_error_tweak(ep, grad, 0), math.sqrt(24.0 / 8.0))
    # fewer than 32 bins at every level: the "reliable level" falls back to the unbinned one
    assert ep.reliable_level() == 0
    # with 64 * 32 values the highest level that still has >= 32 bins is level 6 (2048 / 2^6 = 32)
    big = ErrorPropagator(2)
    for x in range(2048):
        big.push(float(x % 7), float((x % 7) ** 2))
    assert big.count[6] == 32 and big.count[7] == 16 and big.reliable_level() == 6
