"""MonteCarlo drivers — host-side mirror of src/monte_carlo.jl and src/metropolis.jl:4-29.

The schedules (what runs when) follow the reference line by line; the sweeps themselves run on the
GPU through libcsmc.  Differences that are deliberate and documented in DESIGN.md:
  * sweeps visit every site once per sweep in colour order (race-free colour passes);
  * parallel tempering keeps configurations where they are and exchanges temperatures; results are
    attributed to temperature slots, so files / observables per temperature read the same;
  * ``report_interval`` / ``checkpoint_rate`` of 0 mean "never" (the reference divides by them,
    src/monte_carlo.jl:355,383);
  * ``mc.T`` may be a sequence: several temperature slots per process (one process per GPU).
"""
from __future__ import annotations

import datetime
import math
import os
import warnings
from collections import namedtuple

import numpy as np

from . import hdf5 as h5
from . import parallel
from .metropolis import Metropolis, SweepAlgorithm
from .observables import Observables

SimulationParameters = namedtuple(
    "SimulationParameters",
    ["t_thermalization", "t_deterministic", "t_measurement", "probe_rate", "swap_rate",
     "overrelaxation_rate", "report_interval", "checkpoint_rate"])          # src/metropolis.jl:4-13


def MCParamsBuffer(d: dict) -> SimulationParameters:
    """src/monte_carlo.jl:10-27 (fills defaults into the caller's dict, warns on unknown keys)."""
    allowed = list(SimulationParameters._fields)
    defaults = [1, 1, 1, 1, 1, 10, 0, 0]
    vals = []
    for k, dv in zip(allowed, defaults):
        if k not in d:
            d[k] = dv
        vals.append(int(d[k]))
    for k in d:
        if k not in allowed:
            warnings.warn(f"'{k}' not a valid MC parameter; ignoring")
    return SimulationParameters(*vals)


class MonteCarlo:
    """src/monte_carlo.jl:56-118.  Fields as src/metropolis.jl:15-29."""

    def __init__(self, T, lattice, parameters: dict, constraint=lambda x: 0.0, weight: float = 0.0,
                 outpath: str = "", outprefix: str = "configuration", inparams: dict | None = None,
                 overwrite: bool = True, sigma0: float = 60, corr: bool = False, ks=None, seed: int | None = None,
                 device: int | None = None):
        ks = None if ks is None else np.asarray(ks, dtype=np.float64)
        if corr and (ks is None or ks.size == 0):
            raise ValueError("No momentum vectors provided for correlation calculations!")        # src/monte_carlo.jl:64-65
        if not corr and ks is not None and ks.size != 0:
            warnings.warn("Momentum vectors provided but correlation calculations not requested!")  # :66-67
        self.temperatures = np.atleast_1d(np.asarray(T, dtype=np.float64)).copy()
        self.T = float(self.temperatures[0])
        self.observables = Observables(0)
        self.lattice = lattice.copy()                                   # deepcopy(lattice), :74
        self.parameters = MCParamsBuffer(parameters)
        self.lambda_ = 0.0
        self.weight = weight
        self.constraint = constraint
        self.sigma = sigma0
        self.sigma0 = sigma0
        self.corr = corr
        self.momentum_vectors = ks
        self.seed = int(np.random.SeedSequence().entropy % (1 << 63)) if seed is None else int(seed)
        self.device = device
        self._engine = None
        self.replica_spins = [self.lattice.spins] + [lattice.copy().spins for _ in range(len(self.temperatures) - 1)]
        self.observables_all = [self.observables] + [Observables(0) for _ in range(len(self.temperatures) - 1)]
        self.statistics = {}

        rank, comm_size = parallel.comm_info()                          # :85-92
        self.rank, self.comm_size = rank, comm_size
        self.outdir, self.outprefix, self.paramsfile = outpath, outprefix, None
        # temperature slots: the reference has one per MPI rank and names its files by rank; with several
        # slots per process the files are named by global slot (== rank when there is one slot per process)
        if comm_size > 1:
            _, self.slot_base, _ = parallel.gather_temperatures(self.temperatures)
            # the exchange decisions are drawn redundantly on every rank from the shared Philox stream: all
            # processes of a job must use one seed (rank 0's when none was given)
            seeds = parallel.allgather_objects(self.seed)
            if seed is None:
                self.seed = int(seeds[0])
            elif len(set(seeds)) != 1:
                raise ValueError("every process of a parallel-tempering job must be given the same seed")
        else:
            self.slot_base = 0
        if len(outpath) > 0:                                            # :95-113
            if rank == 0 and not os.path.isdir(outpath):
                os.makedirs(outpath, exist_ok=True)
            parallel.barrier()
            self.outpath = outpath + f"{outprefix}_{self.slot_base}.h5"
            self.paramsfile = outpath + outprefix + ".h5.params"
            if rank == 0 and not os.path.isfile(self.paramsfile) and overwrite:
                h5.create_params_file(self, self.paramsfile)
                if inparams:
                    h5.write_attributes(self.paramsfile, inparams)
            parallel.barrier()
            for r, T_r in enumerate(self.temperatures):
                path = outpath + f"{outprefix}_{self.slot_base + r}.h5"
                if not os.path.isfile(path) and overwrite:
                    print(f"Creating new file {os.path.basename(path)} for output on rank {rank}")
                    h5.initialize_hdf5(self, self.paramsfile, outpath=path, T=T_r, spins=self.replica_spins[r])
        else:
            self.outpath = outpath

    # ---- device state ------------------------------------------------------------------------------
    def _device(self, n_replicas=None, replica_base=None):
        n_replicas = len(self.temperatures) if n_replicas is None else n_replicas
        replica_base = self.slot_base if replica_base is None else replica_base   # global index of local replica 0
        if self._engine is None or self._engine.n_replicas != n_replicas or self._engine.replica_base != replica_base:
            from . import _lib
            if self._engine is not None:
                self._engine.close()
            dev = self.device
            if dev is None:
                dev = int(os.environ.get("LOCAL_RANK", "0"))
            self._engine = _lib.Engine(self.lattice._model, n_replicas=n_replicas, seed=self.seed, device=dev,
                                       replica_base=replica_base)
        return self._engine

    def _upload(self):
        eng = self._device()
        self.replica_spins[0] = self.lattice.spins
        for r, s in enumerate(self.replica_spins):
            v = s.T
            eng.set_spins(v if v.flags["C_CONTIGUOUS"] else np.ascontiguousarray(v), replica=r)
        return eng

    def _download(self):
        eng = self._engine
        for r in range(len(self.replica_spins)):
            self.replica_spins[r] = np.asfortranarray(eng.get_spins(replica=r).T)
        self.lattice.spins = self.replica_spins[0]


def _print_acceptance(T, R, accept_total):
    print(f"Acceptance rate at T={T}: {round(R / accept_total * 100, 5)} % ")


def simulated_annealing(mc: MonteCarlo, schedule, T0: float = 1.0, alg=None):
    """src/monte_carlo.jl:157-190.  One device call per temperature (csmc_anneal_temperature[_cone]) for
    Metropolis(), MetropolisAdaptive() and MetropolisFixedCone(); any other callable is driven sweep by
    sweep through the ``alg(mc, T)`` seam."""
    alg = Metropolis() if alg is None else alg
    p = mc.parameters
    T = T0
    time = 1
    out = len(mc.outpath) > 0
    accept_total = p.t_thermalization * mc.lattice.size                         # :163-166
    if p.overrelaxation_rate != 0:
        accept_total /= p.overrelaxation_rate
    eng = mc._upload()
    kind = alg.kind if isinstance(alg, SweepAlgorithm) else None
    while T > mc.T:                                                             # :168
        R = 0.0
        mc.sigma = mc.sigma0                                                    # :171
        if kind == "metropolis":
            R = float(eng.anneal_temperature(T, p.t_thermalization, p.overrelaxation_rate)[0])
        elif kind in ("adaptive", "fixed_cone"):
            acc, sig = eng.anneal_temperature_cone(T, mc.sigma, kind == "adaptive", p.t_thermalization, p.overrelaxation_rate)
            R, mc.sigma = float(acc[0]), float(sig[0])
        else:
            # sweep by sweep through the alg(mc, T) seam.  The library's own algorithms work on the device copy
            # (_device_resident); any other callable sees and may edit mc.lattice.spins, so the state is brought to
            # the host before the call and back to the device after it.
            own = isinstance(alg, SweepAlgorithm)

            def call_alg():
                if own:
                    return alg(mc, T)
                mc._download()
                acc = alg(mc, T)
                mc._upload()
                return acc
            mc._device_resident = own
            try:
                t = 1
                while t < p.t_thermalization:                                   # :172-182
                    if p.overrelaxation_rate != 0:
                        eng.overrelax(1)
                        if t % p.overrelaxation_rate == 0:
                            R += call_alg()
                    else:
                        R += call_alg()
                    t += 1
            finally:
                mc._device_resident = False
        _print_acceptance(T, R, accept_total)                                   # :183
        T = schedule(time)                                                      # :184
        time += 1
        if out:                                                                 # :186-188
            mc._download()
            h5.write_MC_checkpoint(mc)
    mc._download()


def deterministic_updates(mc: MonteCarlo):
    """src/monte_carlo.jl:201-213.  The reference performs t_deterministic - 1 single-site updates at
    random sites; the device version performs ceil((t_deterministic - 1) / N) colour-ordered full
    sweeps of the same update (at least as many site updates, every site visited)."""
    n_updates = max(mc.parameters.t_deterministic - 1, 0)
    n_sweeps = int(math.ceil(n_updates / mc.lattice.size)) if n_updates else 0
    eng = mc._upload()
    done = 0
    while done < n_sweeps:
        k = min(4096, n_sweeps - done)
        eng.deterministic(k)
        done += k
    mc._download()


def print_runtime_statistics(mc, t, stats, T_all, n_local_base):
    """src/helper.jl:25-78 (rank 0 prints; per-slot rates come from the device counters)."""
    p = mc.parameters
    total_sweeps = p.t_thermalization + p.t_measurement
    acc, exch = stats["accepted_local"], stats["exchanges"]
    dt = t - stats["t_prev"]
    rate = p.overrelaxation_rate if p.overrelaxation_rate != 0 else 1
    attempted_local = dt * mc.lattice.size / rate                                   # :30
    n = len(T_all)
    if mc.rank == 0:
        s = f"Sweep {t} / {total_sweeps} ({100.0 * t / total_sweeps:.1f}%)\n"
        s += f"\t\tthermalized : {'YES' if t >= p.t_thermalization else 'NO'}\n"
        for k in range(n):
            a = (acc[k] - stats["acc_prev"][k]) / attempted_local * 100.0
            if n == 1:
                s += f"\t\tupdate acceptance rate : {a:.2f}%\tsigma : {mc.sigma:.2f}\n"
            else:
                att = dt / p.swap_rate / (2.0 if k in (0, n - 1) else 1.0)             # :39
                e = (exch[k] - stats["exch_prev"][k]) / att * 100.0 if att > 0 else 0.0
                s += f"\t\tsimulation {k} update acceptance rate : {a:.2f}%\tsigma : {mc.sigma:.2f}\n"
                s += f"\t\tsimulation {k} replica exchange acceptance rate : {e:.2f}%\n"
        print(s + "\n", end="")
    stats["acc_prev"], stats["exch_prev"], stats["t_prev"] = acc.copy(), exch.copy(), t


def parallel_tempering(mc: MonteCarlo, saveIC=(), alg=None):
    """src/monte_carlo.jl:235-398.

    Temperature slots: the concatenation over ranks of each process's ``mc.temperatures`` (one slot
    per MPI rank in the reference).  The device loop (csmc_pt_run) is chunked at checkpoint / report
    boundaries, which are the only points where the host needs the state."""
    alg = Metropolis() if alg is None else alg
    kinds = {"metropolis": 0, "adaptive": 1, "fixed_cone": 2}
    if not (isinstance(alg, SweepAlgorithm) and alg.kind in kinds):
        raise NotImplementedError("parallel_tempering on the device supports alg=Metropolis(), MetropolisAdaptive() "
                                  "and MetropolisFixedCone()")
    p = mc.parameters
    rank, comm_size = parallel.comm_info()
    out = len(mc.outpath) > 0
    T_all, base, counts = parallel.gather_temperatures(mc.temperatures)             # :246-256
    n_slots, R = len(T_all), len(mc.temperatures)
    if n_slots == 1:
        warnings.warn("a single temperature slot; no replica exchanges will occur!")  # :258

    from . import _lib
    mc._device(n_replicas=R, replica_base=base)
    eng = mc._upload()
    if comm_size > 1:
        uid = parallel.broadcast_unique_id(_lib.comm_unique_id)
        eng.comm_init(comm_size, rank, uid)
    eng.pt_init(T_all)                                                               # E = total_energy, :265
    eng.set_sigma(mc.sigma)                                                          # mc.sigma = sigma0 on every slot
    if mc.corr:                                                                      # :371-375
        eng.pt_set_momenta(mc.lattice.unit_cell.lattice_vectors, mc.lattice.unit_cell.basis, mc.momentum_vectors)

    saveIC = [int(s) for s in saveIC]
    path = os.path.dirname(mc.outpath)
    if saveIC and out:                                                               # :278-283
        for s in saveIC:
            d = os.path.join(path, f"IC_{s}")
            if rank == 0 and not os.path.isdir(d):
                print(f"Initializing IC collection on rank {s}")
                os.makedirs(d, exist_ok=True)
        parallel.barrier()

    def slot_paths(slot):
        return os.path.join(mc.outdir, f"{mc.outprefix}_{slot}.h5") if out else None

    if out:  # the configuration files (one per slot) were created by MonteCarlo(); wait for every rank's
        parallel.barrier()

    if rank == 0:
        print("Running sweeps on %s." % datetime.datetime.now().strftime("%d %b %Y %H:%M:%S"))  # :286

    total = p.t_thermalization + p.t_measurement                                     # :276
    params = p._asdict()
    params["algorithm"] = kinds[alg.kind]
    stats = {"acc_prev": np.zeros(n_slots), "exch_prev": np.zeros(n_slots), "t_prev": 0}

    def checkpoint(sweep):
        """:355-364 — every slot's configuration is written by the process that holds it."""
        slots = eng.pt_slots()
        for r in range(R):
            slot = int(slots[base + r])
            spins = np.asfortranarray(eng.get_spins(replica=r).T)
            if out:
                h5.write_MC_checkpoint(mc, outpath=slot_paths(slot), spins=spins)
                if slot in saveIC:
                    timestep = (sweep - p.t_thermalization) // p.checkpoint_rate
                    h5.write_initial_configuration(os.path.join(path, f"IC_{slot}", f"IC_{timestep}.h5"), mc,
                                                   spins=spins, T=T_all[slot], config_path=slot_paths(slot))

    # host events: a checkpoint happens *inside* iteration `sweep` (after its sweeps, :353-365), a
    # report after `sweep` has been incremented (:380-387).  Chunks end right after such iterations.
    sweep = 0
    while sweep < total:
        nxt = total
        if p.checkpoint_rate > 0:
            c = max(sweep, p.t_thermalization)
            c = ((c + p.checkpoint_rate - 1) // p.checkpoint_rate) * p.checkpoint_rate
            if c < total:
                nxt = min(nxt, c + 1)
        if p.report_interval > 0:
            nxt = min(nxt, ((sweep // p.report_interval) + 1) * p.report_interval)
        eng.pt_run(params, sweep, nxt)
        last = nxt - 1
        if p.checkpoint_rate > 0 and last >= p.t_thermalization and last % p.checkpoint_rate == 0 and (out or saveIC):
            checkpoint(last)
        sweep = nxt
        if p.report_interval > 0 and sweep % p.report_interval == 0:
            a, e = eng.pt_stats()
            stats["accepted_local"], stats["exchanges"] = a, e
            print_runtime_statistics(mc, sweep, stats, T_all, base)

    # ---- results, attributed to temperature slots ------------------------------------------------------
    E, M = eng.pt_series()
    slots = eng.pt_slots()
    acc, exch = eng.pt_stats()
    sig = eng.get_sigma()            # cone widths of the local replicas (they travel with the temperature slot)
    mc.sigma_all = {int(slots[base + r]): float(sig[r]) for r in range(R)}
    mc.sigma = mc.sigma_all.get(base, float(sig[0]))
    mc.statistics = {"accepted_local": acc, "exchanges": exch, "slot_of_replica": slots, "sigma": mc.sigma_all,
                     "temperatures": T_all, "energy_series": E, "magnetization_series": M}
    for r in range(R):                                                               # update_observables!, :368-370
        obs = mc.observables_all[r]
        for k in range(E.shape[0]):
            obs.energy.push(E[k, base + r], E[k, base + r] ** 2)
            obs.magnetization.push(M[k, base + r], M[k, base + r] ** 2)
    if mc.corr:   # mean(SSF) per temperature slot: sum the ranks' contributions, divide by the probe count
        sums, n_ssf = eng.pt_ssf()
        total = sum(parallel.allgather_objects(sums))
        for r in range(R):
            mc.observables_all[r].correlations = total[base + r] / max(n_ssf, 1)
    # configurations by slot: the reference leaves in mc.lattice.spins the configuration that sits at
    # temperature mc.T at the end (it swapped configurations between ranks, :336-347)
    local_cfg = [np.asfortranarray(eng.get_spins(replica=r).T) for r in range(R)]
    by_slot = parallel.collect_by_slot(local_cfg, slots[base:base + R], n_slots)
    for r in range(R):
        mc.replica_spins[r] = by_slot[base + r]
    mc.lattice.spins = mc.replica_spins[0]

    if out:                                                                          # :390-394
        if rank == 0:
            print("Writing observables on %s." % datetime.datetime.now().strftime("%d %b %Y %H:%M:%S"))
        for r in range(R):
            h5.write_final_observables(mc, outpath=slot_paths(base + r), spins=mc.replica_spins[r],
                                       observables=mc.observables_all[r], T=T_all[base + r])
            if mc.corr:                                                              # src/hdf5.jl:229-236
                f = h5._open(slot_paths(base + r), "r+")
                h5.overwrite_keys(f, {"spin_correlations/SSF": mc.observables_all[r].correlations,
                                      "spin_correlations/SSF_momentum": mc.momentum_vectors})
                f.close()
    if rank == 0:
        print("Simulation finished on %s." % datetime.datetime.now().strftime("%d %b %Y %H:%M:%S"))   # :396
    return
