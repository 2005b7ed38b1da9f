"""Statistical parity (BASELINE.json north_star, correctness part 3): thermal averages from the GPU
Markov chain agree with exact results and with the oracle's CPU run of the reference algorithm within
their statistical error bars.  The chains differ by construction (colour-ordered sweeps + Philox on
the GPU, random sites with replacement + xoshiro in the reference), so agreement is in distribution.
Error bars come from a binning analysis of the measured series; thresholds are 5 sigma."""
import math

import numpy as np
import pytest

import classicalspinmc.jl_b200 as csm
from classicalspinmc.jl_b200 import _lib
from classicalspinmc.jl_b200._abi import FLAG_JIT, ModelData
from oracle import oracle as orc
from tests import models

pytestmark = pytest.mark.gpu


def binned_error(x, nbins=32):
    x = np.asarray(x, dtype=float)
    n = len(x) // nbins * nbins
    b = x[:n].reshape(nbins, -1).mean(axis=1)
    return b.mean(), b.std(ddof=1) / math.sqrt(nbins)


def langevin(x):
    return 1.0 / math.tanh(x) - 1.0 / x


@pytest.mark.parametrize("kind", ["uniform", "cone"])
def test_free_spins_follow_the_langevin_function(kind):
    """Non-interacting spins in a field: <s_z> = S L(h S / T) exactly."""
    h, S, T = 0.7, 1.0, 0.5
    uc = csm.Square()
    csm.addZeemanCoupling(uc, 1, np.array([0.0, 0.0, h]))
    md = ModelData(uc, (64, 64), S)
    eng = _lib.Engine(md, seed=11, flags=FLAG_JIT)
    eng.randomize(3)
    run = (lambda n: eng.metropolis(T, n)) if kind == "uniform" else (lambda n: eng.metropolis_cone(T, 1.5, False, n)[0])
    run(100)
    mz = []
    for _ in range(320):
        run(2)
        mz.append(eng.magnetization_vector()[0][2] / md.n_sites)
    mean, err = binned_error(mz)
    exact = S * langevin(h * S / T)
    assert abs(mean - exact) < 5 * err + 1e-4, (mean, exact, err)


def test_open_heisenberg_chain_matches_fisher():
    """Open classical Heisenberg chain (Fisher 1964): E / bond = -|J| S^2 L(|J| S^2 / T)."""
    J, S, Lc = -1.0, 1.0, 64
    md = ModelData(models.chain_heisenberg(J), (Lc,), S, "open")
    Ts = np.array([0.3, 0.6, 1.2, 2.4])
    R = 64                      # 16 independent chains per temperature
    T_all = np.repeat(Ts, R // len(Ts))
    eng = _lib.Engine(md, n_replicas=R, seed=5)
    eng.randomize(9)
    eng.set_temperatures(T_all)
    eng.cycles_async(300, 2, 1)
    series = []
    for _ in range(160):
        eng.cycles_async(5, 2, 1)
        series.append(eng.total_energy() / (Lc - 1))
    series = np.array(series)
    for k, T in enumerate(Ts):
        cols = series[:, k * 16:(k + 1) * 16].mean(axis=1)
        mean, err = binned_error(cols, 16)
        exact = -abs(J) * S * S * langevin(abs(J) * S * S / T)
        assert abs(mean - exact) < 5 * err + 2e-4, (T, mean, exact, err)


def test_thermal_averages_match_reference_algorithm():
    """Kitaev-Gamma honeycomb L=6 at T=0.4: E/N, |M|/N, specific heat and Metropolis acceptance from the
    GPU chain vs the oracle's run of the reference algorithm (random-site Metropolis + sequential
    overrelaxation), both through the reference's parallel_tempering loop with one temperature."""
    md = ModelData(models.kitaev_honeycomb(), (6, 6), 1.0)
    N = md.n_sites
    T = 0.4
    p = dict(t_thermalization=2000, t_measurement=40000, probe_rate=10, swap_rate=10 ** 9, overrelaxation_rate=5)
    lat = orc.OracleLattice(md)
    s = lat.randomize(seed=1)
    E_ref, M_ref, acc_ref, _ = lat.parallel_tempering(s, [T], p["t_thermalization"], p["t_measurement"],
                                                      p["probe_rate"], p["swap_rate"], p["overrelaxation_rate"], seed=77)
    eng = _lib.Engine(md, n_replicas=1, seed=99)
    eng.set_spins(lat.randomize(seed=2))
    eng.pt_init([T])
    eng.pt_run(p, 0, p["t_thermalization"] + p["t_measurement"])
    E, M = eng.pt_series()
    acc, _ = eng.pt_stats()
    assert E.shape == E_ref.shape == (p["t_measurement"] // p["probe_rate"], 1)
    for a, b, scale in ((E[:, 0] / N, E_ref[:, 0] / N, 1), (M[:, 0] / N, M_ref[:, 0] / N, 1)):
        ma, ea = binned_error(a)
        mb, eb = binned_error(b)
        assert abs(ma - mb) < 5 * math.hypot(ea, eb), (ma, mb, ea, eb)
    # specific heat c = (<E^2> - <E>^2) / (T^2 N), src/observables.jl:42; error from bin-to-bin scatter
    def cv(x):
        bins = x[: len(x) // 32 * 32].reshape(32, -1)
        c = bins.var(axis=1) / (T * T * N)
        return c.mean(), c.std(ddof=1) / math.sqrt(32)
    ca, ea = cv(E[:, 0])
    cb, eb = cv(E_ref[:, 0])
    assert abs(ca - cb) < 5 * math.hypot(ea, eb), (ca, cb, ea, eb)
    # acceptance rate of uniform proposals is a thermal average too
    n_metro = (p["t_thermalization"] + p["t_measurement"]) // p["overrelaxation_rate"]
    ra, rb = acc[0] / (n_metro * N), acc_ref[0] / (n_metro * N)
    assert abs(ra - rb) < 0.01, (ra, rb)


def test_parallel_tempering_matches_reference_algorithm():
    """Replica exchange: per-temperature <E> and exchange acceptance of the device loop (temperatures
    swapped) vs the oracle's reference loop (configurations swapped, src/monte_carlo.jl:308-349)."""
    md = ModelData(models.kitaev_honeycomb(), (4, 4), 1.0)
    N = md.n_sites
    Ts = np.geomspace(0.15, 1.0, 6)
    R = len(Ts)
    p = dict(t_thermalization=2000, t_measurement=30000, probe_rate=10, swap_rate=10, overrelaxation_rate=5)
    lat = orc.OracleLattice(md)
    spins = np.concatenate([lat.randomize(seed=10 + r) for r in range(R)])
    E_ref, M_ref, acc_ref, ex_ref = lat.parallel_tempering(spins, Ts, p["t_thermalization"], p["t_measurement"],
                                                           p["probe_rate"], p["swap_rate"], p["overrelaxation_rate"], seed=5)
    eng = _lib.Engine(md, n_replicas=R, seed=1234)
    for r in range(R):
        eng.set_spins(lat.randomize(seed=40 + r), replica=r)
    eng.pt_init(Ts)
    total = p["t_thermalization"] + p["t_measurement"]
    eng.pt_run(p, 0, total // 2)
    eng.pt_run(p, total // 2, total)           # chunked runs continue the same loop
    E, M = eng.pt_series()
    acc, ex = eng.pt_stats()
    assert E.shape == E_ref.shape
    for k in range(R):
        ma, ea = binned_error(E[:, k] / N)
        mb, eb = binned_error(E_ref[:, k] / N)
        assert abs(ma - mb) < 5 * math.hypot(ea, eb) + 1e-4, (k, ma, mb, ea, eb)
    n_attempt = total // p["swap_rate"]
    rate, rate_ref = ex / n_attempt, ex_ref / n_attempt
    assert np.all(np.abs(rate - rate_ref) < 0.05), (rate, rate_ref)
    assert sorted(eng.pt_slots().tolist()) == list(range(R))
    assert ex.sum() > 0 and np.all(acc > 0)


def test_resident_kernel_parallel_tempering_matches_pass_kernels():
    """Small lattices run whole sweep schedules in one launch per replica block (resident kernel).  Over a
    short PT run (before chaotic amplification of ulp-level differences in FMA contraction) the series,
    exchange decisions and configurations agree with the per-colour pass kernels to 1e-9; over a long
    run the thermal averages agree statistically."""
    from classicalspinmc.jl_b200._abi import FLAG_NO_RESIDENT
    md = ModelData(models.pyrochlore_local(), (4, 4, 4), 0.5)
    lat = orc.OracleLattice(md)
    Ts = np.geomspace(0.02, 0.5, 6)
    p = dict(t_thermalization=0, t_measurement=16, probe_rate=1, swap_rate=2, overrelaxation_rate=3)
    res = []
    for flags in (FLAG_JIT, FLAG_JIT | FLAG_NO_RESIDENT):
        eng = _lib.Engine(md, n_replicas=6, seed=77, flags=flags)
        for r in range(6):
            eng.set_spins(lat.randomize(seed=200 + r), replica=r)
        eng.pt_init(Ts)
        eng.pt_run(p, 0, 7)
        eng.pt_run(p, 7, 16)
        E, M = eng.pt_series()
        res.append((E, M, eng.pt_slots(), eng.pt_stats(), [eng.get_spins(r) for r in range(6)], eng.kernel_mode, eng.launches))
    assert res[0][5] == 3 and res[1][5] == 2
    assert res[0][0].shape == (16, 6)
    assert np.allclose(res[0][0], res[1][0], rtol=1e-9, atol=1e-9) and np.allclose(res[0][1], res[1][1], rtol=1e-9, atol=1e-9)
    assert np.array_equal(res[0][2], res[1][2])
    assert np.array_equal(res[0][3][0], res[1][3][0]) and np.array_equal(res[0][3][1], res[1][3][1])
    for a, b in zip(res[0][4], res[1][4]):
        assert np.abs(a - b).max() <= 1e-9
    assert res[0][6] < res[1][6] / 3          # far fewer launches
    # long run: statistical agreement of <E> per temperature slot
    p = dict(t_thermalization=500, t_measurement=6000, probe_rate=5, swap_rate=10, overrelaxation_rate=5)
    means = []
    for flags in (FLAG_JIT, FLAG_JIT | FLAG_NO_RESIDENT):
        eng = _lib.Engine(md, n_replicas=6, seed=5 + flags, flags=flags)
        eng.randomize(11 + flags)
        eng.pt_init(Ts)
        eng.pt_run(p, 0, 6500)
        E, _ = eng.pt_series()
        means.append([binned_error(E[:, k] / md.n_sites) for k in range(6)])
    for (ma, ea), (mb, eb) in zip(*means):
        assert abs(ma - mb) < 5 * math.hypot(ea, eb) + 1e-5, (ma, mb, ea, eb)


@pytest.mark.parametrize("flags", [0, FLAG_JIT], ids=["pass-kernels", "resident"])
def test_parallel_tempering_with_adaptive_cone_moves(flags):
    """parallel_tempering!(mc; alg=MetropolisAdaptive()): cone widths adapt per temperature slot (small at
    low T, saturating at the clamp at high T), travel with the slot on exchanges, and the thermal averages
    agree with plain Metropolis PT."""
    from classicalspinmc.jl_b200._abi import FLAG_NO_RESIDENT
    md = ModelData(models.kitaev_honeycomb(), (4, 4), 1.0)
    N = md.n_sites
    Ts = np.geomspace(0.05, 1.5, 6)
    p = dict(t_thermalization=1500, t_measurement=15000, probe_rate=5, swap_rate=10, overrelaxation_rate=5)
    out = {}
    for alg in (0, 1):
        eng = _lib.Engine(md, n_replicas=6, seed=321 + alg, flags=(flags or FLAG_NO_RESIDENT))
        eng.randomize(17 + alg)
        eng.pt_init(Ts)
        eng.set_sigma(60.0)
        eng.pt_run(dict(p, algorithm=alg), 0, 16500)
        E, _ = eng.pt_series()
        slots = eng.pt_slots()
        sig = eng.get_sigma()
        out[alg] = ([binned_error(E[:, k] / N) for k in range(6)], {int(slots[r]): sig[r] for r in range(6)}, eng.pt_stats())
    for (ma, ea), (mb, eb) in zip(out[0][0], out[1][0]):
        assert abs(ma - mb) < 5 * math.hypot(ea, eb) + 1e-4, (ma, mb, ea, eb)
    sigma_by_slot = out[1][1]
    assert sigma_by_slot[0] < 2.0 < sigma_by_slot[5]            # narrow cone when cold, wide when hot
    assert all(0.0 <= v <= 100.0 for v in sigma_by_slot.values())
    acc_plain, acc_adapt = out[0][2][0], out[1][2][0]
    assert acc_adapt[0] > 2 * acc_plain[0]                       # adaptive cone: far higher acceptance at low T
    assert np.all(out[1][1 + 1][1] > 0)                          # exchanges happened


def test_annealing_reaches_the_reference_ground_state_energy():
    """test/mctests.jl:43-50 through the host mirror: simulated_annealing! + deterministic_updates! on
    the Kitaev-Gamma honeycomb, round(E/N, digits=4) == -0.6444."""
    uc = models.kitaev_honeycomb()
    lat = csm.Lattice((4, 4), uc, 1, rng=np.random.default_rng(5))
    T0, T = 1.0, 1e-7
    params = {"t_thermalization": int(1e4), "overrelaxation_rate": 10, "t_deterministic": int(1e6)}
    mc = csm.MonteCarlo(T, lat, params, seed=2)
    before = lat.spins.copy()
    csm.simulated_annealing(mc, lambda x: T0 * 0.9 ** x, T0)
    csm.deterministic_updates(mc)
    E = csm.energy_density(mc.lattice)
    assert round(E, 4) == -0.6444
    assert np.array_equal(lat.spins, before)        # the user's lattice is not mutated (deepcopy, :74)
    assert np.allclose(np.linalg.norm(mc.lattice.spins, axis=0), 1.0, atol=1e-9)


def test_readme_example_square_lattice():
    """README.md:28-88 (C1): square Heisenberg L=4, J=-I, h=0.1 z, annealing + deterministic updates:
    the ferromagnet aligned with the field, E/N = -2 - 0.1 exactly."""
    uc = models.square_heisenberg(J=-1.0, h=(0.0, 0.0, 0.1))
    lat = csm.Lattice((4, 4), uc, 1.0, bc="periodic", rng=np.random.default_rng(1))
    mc = csm.MonteCarlo(1e-7, lat, {"t_thermalization": int(1e4), "t_deterministic": int(1e5), "overrelaxation_rate": 10}, seed=3)
    csm.simulated_annealing(mc, lambda x: 1.0 * 0.9 ** x, 1.0)
    csm.deterministic_updates(mc)
    assert abs(csm.energy_density(mc.lattice) - (-2.1)) < 1e-6
    assert abs(csm.get_magnetization(mc.lattice) - 16.0) < 1e-5
    assert np.allclose(mc.lattice.spins[2], 1.0, atol=1e-5)
