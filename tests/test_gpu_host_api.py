"""The reference's own tests restated against the host mirror (Python stand-in for the Julia layer),
running on the GPU through libcsmc, plus driver-level behaviour: output files, parallel tempering
through `parallel_tempering`, error reporting across the C-ABI."""
import ctypes
import os

import numpy as np
import pytest

import classicalspinmc.jl_b200 as csm
from classicalspinmc.jl_b200 import _lib
from classicalspinmc.jl_b200 import hdf5 as h5
from classicalspinmc.jl_b200._abi import FLAG_FORCE_GENERIC, ModelData
from tests import models

pytestmark = pytest.mark.gpu


def test_lattice_tests_jl():
    # test/latticetests.jl:3-31, line by line
    U = csm.Square()
    lat = csm.Lattice((2, 2), U, 1.0)
    assert np.all(np.round(np.linalg.norm(lat.spins, axis=0), 9) == 1.0)

    U = csm.Square()
    h = np.array([1.0, 0.0, 0.0])
    csm.addZeemanCoupling(U, 1, h)
    lat = csm.Lattice((1, 1), U, 1.0)
    lat.spins[:] = np.array([1.0, 0.0, 0.0])[:, None]
    assert -1.0 == csm.total_energy(lat)
    assert tuple(-h) == csm.get_local_field(lat, 1)

    U = csm.Square()
    J = -1.0 * np.eye(3)
    csm.addBilinear(U, 1, 1, J, (1, 0))
    csm.addBilinear(U, 1, 1, J, (-1, 0))
    csm.addBilinear(U, 1, 1, J, (0, 1))
    csm.addBilinear(U, 1, 1, J, (0, -1))
    lat = csm.Lattice((2, 2), U, 1.0)
    lat.spins[:] = np.array([1.0, 0.0, 0.0])[:, None]
    assert -2.0 == csm.total_energy(lat) / lat.size
    # the host array is the state of a bare Lattice: mutate and re-evaluate
    lat.spins[:, 0] = (0.0, 0.0, 1.0)
    assert abs(csm.total_energy(lat) - (-4.0)) < 1e-14
    assert abs(csm.get_magnetization(lat) - np.linalg.norm([3.0, 0.0, 1.0])) < 1e-14


def test_adaptive_annealing_mctests_jl():
    # test/mctests.jl:52-58: MetropolisAdaptive, no deterministic step
    lat = csm.Lattice((4, 4), models.kitaev_honeycomb(), 1, rng=np.random.default_rng(8))
    params = {"t_thermalization": int(1e4), "overrelaxation_rate": 10, "t_deterministic": int(1e6)}
    mc = csm.MonteCarlo(1e-7, lat, params, seed=4)
    csm.simulated_annealing(mc, lambda x: 1.0 * 0.9 ** x, 1.0, alg=csm.MetropolisAdaptive())
    assert -0.6444 == round(csm.energy_density(mc.lattice), 4)
    assert 0.0 <= mc.sigma <= 100.0
    # the `alg` seam still accepts any callable (mc, T) -> accepted and drives it sweep by sweep
    calls = []
    def my_alg(mc_, T):
        calls.append(T)
        return csm.Metropolis()(mc_, T)
    mc2 = csm.MonteCarlo(0.5, lat, {"t_thermalization": 21, "overrelaxation_rate": 5}, seed=1)
    csm.simulated_annealing(mc2, lambda x: 1.0 * 0.5 ** x, 1.0, alg=my_alg)
    assert calls == [1.0] * 4


def test_drivers_write_reference_file_layout(tmp_path):
    out = str(tmp_path) + "/"
    lat = csm.Lattice((4, 4), models.kitaev_honeycomb(), 1.0, rng=np.random.default_rng(2))
    params = {"t_thermalization": 50, "overrelaxation_rate": 5}
    mc = csm.MonteCarlo(0.5, lat, params, outpath=out, seed=9)
    csm.simulated_annealing(mc, lambda x: 1.0 * 0.5 ** x, 1.0)      # checkpoint after every temperature
    lat2 = csm.Lattice((4, 4), models.kitaev_honeycomb(), 1.0)
    csm.read_spin_configuration(lat2, out + "configuration_0.h5")
    assert np.array_equal(lat2.spins, mc.lattice.spins)


def test_parallel_tempering_driver_single_process(tmp_path):
    """parallel_tempering! with four temperature slots in one process (one GPU): observables per slot,
    files per slot, configurations attributed to temperatures."""
    out = str(tmp_path) + "/"
    lat = csm.Lattice((4, 4), models.kitaev_honeycomb(), 1.0, rng=np.random.default_rng(3))
    Ts = np.geomspace(0.2, 1.0, 4)
    params = {"t_thermalization": 400, "t_measurement": 1600, "probe_rate": 10, "swap_rate": 10,
              "overrelaxation_rate": 5, "report_interval": 1000, "checkpoint_rate": 800}
    mc = csm.MonteCarlo(Ts, lat, params, outpath=out, seed=12)
    csm.parallel_tempering(mc, saveIC=[0])
    st = mc.statistics
    assert st["energy_series"].shape == (160, 4)
    assert sorted(st["slot_of_replica"].tolist()) == [0, 1, 2, 3]
    assert st["exchanges"].sum() > 0
    E_mean = [o.energy.mean(1) for o in mc.observables_all]
    assert E_mean[0] < E_mean[-1]                       # colder slot, lower energy
    for s in range(4):
        path = out + f"configuration_{s}.h5"
        assert os.path.isfile(path)
        obs = h5.read_observables(path)
        assert set(obs) == {"specific_heat", "specific_heat_err", "susceptibility", "susceptibility_err",
                            "magnetization", "magnetization_err", "energy", "energy_err"}
        assert abs(obs["energy"] - E_mean[s]) < 1e-9
        f = h5._open(path, "r")
        assert abs(h5._get_attr(f, "T") - Ts[s]) < 1e-15
        f.close()
    assert os.path.isfile(out + "IC_0/IC_0.h5") and os.path.isfile(out + "IC_0/IC_1.h5")
    # mc.lattice.spins is the configuration sitting at temperature slot 0 at the end; its energy is the
    # last recorded energy of slot 0 up to the sweeps after the last probe -> just check consistency
    e0 = csm.total_energy(mc.lattice)
    assert np.isfinite(e0) and np.allclose(np.linalg.norm(mc.lattice.spins, axis=0), 1.0, atol=1e-9)


def test_parallel_tempering_driver_with_adaptive_alg():
    lat = csm.Lattice((4, 4), models.kitaev_honeycomb(), 1.0, rng=np.random.default_rng(4))
    params = {"t_thermalization": 300, "t_measurement": 600, "probe_rate": 10, "swap_rate": 10, "overrelaxation_rate": 5}
    mc = csm.MonteCarlo(np.geomspace(0.1, 1.0, 3), lat, params, seed=5)
    csm.parallel_tempering(mc, alg=csm.MetropolisAdaptive())
    assert set(mc.sigma_all) == {0, 1, 2} and mc.sigma_all[0] < mc.sigma_all[2] <= 100.0
    with pytest.raises(NotImplementedError):
        csm.parallel_tempering(mc, alg=csm.MetropolisConstraint())


@pytest.mark.parametrize("case", ["honeycomb", "pyrochlore", "chain"])
def test_structure_factor_matches_reference_formula(case):
    """compute_equal_time_correlations (src/spin_correlations.jl:6-43) vs the oracle's literal restatement."""
    from oracle import oracle as orc
    rng = np.random.default_rng(3)
    if case == "honeycomb":
        uc, shape = models.kitaev_honeycomb(), (6, 5)
    elif case == "pyrochlore":
        uc, shape = models.pyrochlore_local(), (3, 4, 2)
    else:
        uc, shape = models.chain_heisenberg(), (37,)
    lat = csm.Lattice(shape, uc, 1.0, rng=rng)
    ks = rng.uniform(-2 * np.pi, 2 * np.pi, size=(uc.D, 23))
    ks[:, 0] = 0.0
    S = csm.compute_equal_time_correlations(lat, ks)
    o = orc.OracleLattice(lat._model)
    S_ref = o.structure_factor(np.ascontiguousarray(lat.spins.T), lat.site_positions, ks)
    assert S.shape == S_ref.shape == (9, 23)
    assert np.abs(S - S_ref).max() <= 1e-10 * np.abs(S_ref).max()
    # k = 0: S^{uv}(0) = M_u M_v / N
    M = lat.spins.sum(axis=1)
    assert np.allclose(S[:, 0].reshape(3, 3), np.outer(M, M) / lat.size, rtol=1e-12, atol=1e-12)


def test_parallel_tempering_accumulates_structure_factor(tmp_path):
    """mc.corr = true: mean structure factor per temperature slot (src/monte_carlo.jl:371-375); at high T the
    spins are uncorrelated, so trace S(k) ~ S^2 for every k, while the cold slot of the ferromagnet peaks at k=0."""
    uc = models.square_heisenberg(J=-1.0, h=None)
    lat = csm.Lattice((8, 8), uc, 1.0, rng=np.random.default_rng(6))
    ks = np.array([[0.0, np.pi, np.pi / 2], [0.0, np.pi, 0.0]])
    params = {"t_thermalization": 400, "t_measurement": 2000, "probe_rate": 10, "swap_rate": 10, "overrelaxation_rate": 5}
    out = str(tmp_path) + "/"
    mc = csm.MonteCarlo(np.array([0.05, 0.5, 50.0]), lat, params, corr=True, ks=ks, seed=8, outpath=out)
    csm.parallel_tempering(mc)
    S_cold, S_hot = mc.observables_all[0].correlations, mc.observables_all[2].correlations
    assert S_cold.shape == (9, 3)
    tr = lambda S: S[0] + S[4] + S[8]
    assert tr(S_cold)[0] > 0.9 * lat.size and tr(S_cold)[1] < 0.05 * lat.size     # ordered: Bragg peak at k = 0
    assert np.all(np.abs(tr(S_hot) - 1.0) < 0.25)                                  # paramagnet: flat, = S^2
    f = h5._open(out + "configuration_0.h5", "r")
    assert np.allclose(h5._get_jl(f, "spin_correlations/SSF"), S_cold) and np.allclose(h5._get_jl(f, "spin_correlations/SSF_momentum"), ks)
    f.close()


def test_structure_factor_batch_runner_over_configuration_files(tmp_path):
    """runEqualTimeStructureFactor! + compute_equal_time_structure_factor (src/spin_correlations.jl:48-145):
    IC_<n>.h5 files in, per-file spin_correlations group, then the average into a destination file; momenta
    from get_allowed_wavevectors (src/reciprocal.jl:23-29)."""
    uc = models.kitaev_honeycomb()
    lat = csm.Lattice((6, 4), uc, 1.0, rng=np.random.default_rng(1))
    ks = csm.get_allowed_wavevectors(uc, (6, 4))
    mc = csm.MonteCarlo(0.5, lat, {"t_thermalization": 10}, outpath=str(tmp_path) + "/")
    ic_dir = str(tmp_path / "IC_0") + "/"
    os.makedirs(ic_dir)
    rng = np.random.default_rng(2)
    expected = []
    for n in range(3):
        spins = rng.normal(size=(3, lat.size))
        spins /= np.linalg.norm(spins, axis=0)
        h5.write_initial_configuration(ic_dir + f"IC_{n}.h5", mc, spins=spins)
        lat.spins[:, :] = spins
        expected.append(csm.compute_equal_time_correlations(lat, ks))
    csm.runEqualTimeStructureFactor(ic_dir, lat, ks)
    for n in range(3):
        f = h5._open(ic_dir + f"IC_{n}.h5", "r")
        assert np.array_equal(h5._get_jl(f, "spin_correlations/SSF"), expected[n])
        assert np.array_equal(h5._get_jl(f, "spin_correlations/SSF_momentum"), ks)
        assert np.array_equal(np.asarray(h5._get(f, "spins")).T, lat.spins) == (n == 2)
        f.close()
    # already processed files are skipped unless override is set
    f = h5._open(ic_dir + "IC_1.h5", "r+")
    h5.overwrite_keys(f, {"spin_correlations/SSF": np.zeros_like(expected[1])})
    f.close()
    csm.runEqualTimeStructureFactor(ic_dir, lat, ks)
    f = h5._open(ic_dir + "IC_1.h5", "r")
    assert not np.any(h5._get_jl(f, "spin_correlations/SSF"))
    f.close()
    csm.runEqualTimeStructureFactor(ic_dir, lat, ks, True)
    mean = csm.compute_equal_time_structure_factor(ic_dir, mc.outpath)
    assert np.allclose(mean, np.mean(expected, axis=0), rtol=1e-13, atol=1e-13)
    f = h5._open(mc.outpath, "r")
    assert np.allclose(h5._get_jl(f, "spin_correlations/SSF"), mean) and np.array_equal(h5._get_jl(f, "spin_correlations/SSF_momentum"), ks)
    f.close()


def test_example_runners(tmp_path):
    """examples/: the reference's two example scripts (simulated annealing of the Kitaev honeycomb model,
    parallel tempering of the pyrochlore model) on the GPU engine, at reduced sizes."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out1 = str(tmp_path / "sa") + "/"
    r = subprocess.run([sys.executable, os.path.join(root, "examples", "simulated_annealing", "runner.py"), out1,
                        "--t-thermalization", "300", "--t-deterministic", "4000"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-3000:]
    e = float([l for l in r.stdout.splitlines() if l.startswith("energy per site:")][-1].split(":")[1])
    assert -0.9 < e < -0.5                       # Kitaev ferromagnet in a weak field: E/N close to -0.6
    f = h5._open(out1 + "configuration.h5.params", "r")
    assert h5._get_attr(f, "K") == -1.0 and h5._get_attr(f, "t_thermalization") == 300
    assert len(h5._keys(f, "unit_cell/bilinear")) == 3 and h5._has_group(f, "unit_cell/cubic")
    f.close()
    g = h5._open(out1 + "configuration_0.h5", "r")
    spins = np.asarray(h5._get(g, "spins"))
    g.close()
    assert spins.shape == (32, 3) and np.allclose(np.linalg.norm(spins, axis=1), 1.0)
    out2 = str(tmp_path / "pt") + "/"
    r = subprocess.run([sys.executable, os.path.join(root, "examples", "parallel_tempering", "runner.py"), out2,
                        "--temperatures", "6", "--L", "2", "--t-thermalization", "400", "--t-measurement", "2000", "--B", "1.0"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-3000:]
    energies = []
    for slot in range(6):
        obs = h5.read_observables(out2 + f"configuration_{slot}.h5")
        assert {"energy", "energy_err", "specific_heat", "magnetization", "susceptibility"} <= set(obs)
        energies.append(obs["energy"])
    assert energies[0] < energies[-1]            # colder slots sit lower in energy
    assert os.path.isdir(out2 + "IC_0") and len(os.listdir(out2 + "IC_0")) == 2      # checkpoints at sweeps 1000, 2000


def test_errors_cross_the_abi_as_status_codes():
    L = _lib.lib()
    md = ModelData(models.square_heisenberg(), (4, 4), 1.0)
    eng = _lib.Engine(md)
    out = np.zeros(3)
    assert L.csmc_local_field(eng._h, 0, 0, out.ctypes.data_as(ctypes.c_void_p)) == 1          # site is 1-based
    assert b"site out of range" in L.csmc_last_error(eng._h)
    assert L.csmc_local_field(eng._h, 5, 1, out.ctypes.data_as(ctypes.c_void_p)) == 1          # replica
    assert L.csmc_set_spins(eng._h, 0, None) == 1                                              # NULL buffer
    assert L.csmc_total_energy(None, out.ctypes.data_as(ctypes.c_void_p)) == 1                 # NULL handle
    with pytest.raises(_lib.CsmcError, match="temperatures must be > 0"):
        eng.metropolis(0.0, 1)
    with pytest.raises(_lib.CsmcError, match="csmc_pt_init has not been called"):
        eng.pt_run(dict(t_thermalization=1, t_measurement=1, probe_rate=1, swap_rate=1, overrelaxation_rate=1), 0, 1)
    md_bad = ModelData(models.square_heisenberg(), (4, 4), 1.0)
    md_bad.struct.dim = 7
    with pytest.raises(_lib.CsmcError, match="dim must be 1..3"):
        _lib.Engine(md_bad)
    # a handle stays usable after an error
    eng.randomize(1)
    assert np.isfinite(eng.total_energy()[0])


def test_open_boundary_parallel_tempering_generic_kernels():
    """Open boundaries + lattice without a periodic colouring pattern: the explicit-table kernels run the
    whole PT loop; energies stay consistent with a fresh evaluation."""
    md = ModelData(models.square_heisenberg(), (7, 5), 1.0, "open")
    eng = _lib.Engine(md, n_replicas=3, seed=3, flags=FLAG_FORCE_GENERIC)
    assert eng.kernel_mode == 0
    eng.randomize(5)
    eng.pt_init([0.3, 0.6, 1.2])
    p = dict(t_thermalization=100, t_measurement=200, probe_rate=10, swap_rate=5, overrelaxation_rate=4)
    eng.pt_run(p, 0, 300)
    E, M = eng.pt_series()
    assert E.shape == (20, 3) and np.all(np.isfinite(E)) and np.all(M >= 0)
    assert E[:, 0].mean() < E[:, 2].mean()
