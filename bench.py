#!/usr/bin/env python
"""bench.py — single-spin updates/sec of the sweep hot path (BASELINE.json metric).

Workload (config.workload = "C2"): BASELINE.json configs[1], square-lattice Heisenberg L=1024 with a
z-field, single replica per GPU, cycle = 10 overrelaxation sweeps + 1 Metropolis sweep
(checkerboard, 2 colours).  A "step" is `cycles_per_step` such cycles over the lattice.
N > 1 GPUs: the path does not shard a single lattice ("replicas only"), so
every rank runs an independent replica of the same workload (weak scaling, no collective); the
parallel-tempering exchange path over NCCL is measured separately with --workload C3/C4.

    python bench.py --gpus N --steps K --warmup W            # our arm
    python bench.py --impl reference --gpus N --steps K --warmup W   # reference algorithm on host cores
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

B_ALG = {2: 72.0, 4: 120.0}  # algorithmic bytes per single-spin update: 24 (C + 1), SURVEY.md 8(d)


def workload_model(name, L=None):
    from classicalspinmc.jl_b200._abi import ModelData
    from tests import models
    if name == "C2":
        L = L or 1024
        return ModelData(models.square_heisenberg(J=-1.0, h=(0.0, 0.0, 0.1)), (L, L), 1.0), dict(
            workload="C2", lattice="square", L=L, model="Heisenberg J=-1 + h_z=0.1", T=1.0, replicas_per_gpu=1)
    if name == "C3":
        L = L or 256
        return ModelData(models.kitaev_honeycomb(), (L, L), 1.0), dict(
            workload="C3", lattice="honeycomb", L=L, model="Kitaev-Gamma K=-1 G=0.2 Gp=-0.02 h=0.1[111]")
    if name == "C4":
        L = L or 32
        return ModelData(models.pyrochlore_local(), (L, L, L), 0.5), dict(
            workload="C4", lattice="pyrochlore", L=L, model="local-frame Jxx/Jyy/Jzz + Zeeman")
    if name == "C5":
        L = L or 512
        return ModelData(models.triangular_multispin(), (L, L), 1.0), dict(
            workload="C5", lattice="triangular", L=L, model="Heisenberg + cubic + quartic")
    raise SystemExit(f"unknown workload {name}")


class ClockSampler:
    """Samples SM clocks / throttle reasons with NVML while the timed region runs."""

    def __init__(self, index):
        self.samples, self.reasons, self._stop = [], set(), threading.Event()
        self.max_mhz = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        nv = self.nv
        names = {}
        for n in ("HwSlowdown", "HwThermalSlowdown", "SwThermalSlowdown", "SwPowerCap", "HwPowerBrakeSlowdown"):
            v = getattr(nv, "nvmlClocksThrottleReason" + n, None)
            if v is not None:
                names[v] = n
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, n in names.items():
                    if r & bit:
                        self.reasons.add(n)
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        if self.nv:
            self.t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self.nv:
            self.t.join(timeout=2)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        tr = {"HwSlowdown": "hw_slowdown", "HwThermalSlowdown": "hw_thermal_slowdown",
              "SwThermalSlowdown": "sw_thermal_slowdown", "SwPowerCap": "sw_power_cap",
              "HwPowerBrakeSlowdown": "hw_power_brake_slowdown"}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(tr[r] for r in self.reasons)}


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


def ncu_traffic(workload, L=None):
    """dram bytes per launch of the dominant kernel from the committed ncu --set full capture."""
    p = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return d.get(f"{workload}:{L}") if (L and f"{workload}:{L}" in d) else (d.get(workload) if not L or L == 1024 else None)
        except Exception:
            return None
    return None


def cpu_baseline(md, T, or_per_cycle, metro_per_cycle, n_cycles, threads):
    """The oracle's restatement of the reference algorithm (random-site Metropolis with two energy()
    evaluations, sequential overrelaxation) timed on the host cores: `threads` independent replicas."""
    from oracle import oracle as orc
    lat = orc.OracleLattice(md)
    spins = np.concatenate([lat.randomize(seed=12345, replica=r) for r in range(threads)])
    t0 = time.perf_counter()
    updates = lat.cycles(spins, threads, T, n_cycles, or_per_cycle, metro_per_cycle)
    dt = time.perf_counter() - t0
    return updates / dt, updates, dt


def cpu_pt_baseline(md, T_all, sweeps, threads, swap_rate=50, rate=10, seed=3):
    """The oracle's restatement of the reference parallel-tempering loop (src/monte_carlo.jl:289-349: OR every
    sweep, random-site Metropolis + total_energy every `rate`-th, configuration-swapping exchange every
    `swap_rate`-th), one temperature per host thread as examples/parallel_tempering/README.txt:17."""
    from oracle import oracle as orc
    lat = orc.OracleLattice(md)
    R = len(T_all)
    spins = np.concatenate([lat.randomize(seed=12345, replica=r) for r in range(R)])
    t0 = time.perf_counter()
    lat.parallel_tempering(spins, T_all, sweeps, 0, 2000, swap_rate, rate, seed=seed, n_threads=threads)
    dt = time.perf_counter() - t0
    updates = float(sweeps + (sweeps + rate - 1) // rate) * lat.N * R
    return updates / dt, updates, dt


def run_reference(args):
    """Reference arm: the reference's CPU algorithm for the workload (C restatement under oracle/, the
    reference itself is Julia and cannot run here) on all host threads; rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as orc
    md, cfg = workload_model(args.workload, args.L)
    threads = args.cpu_threads or (os.cpu_count() or 1)
    orc.lib()
    pt = args.workload in PT_DEFAULTS
    if pt:
        d = PT_DEFAULTS[args.workload]
        R = args.replicas or d["R"]
        T_all = np.geomspace(d["Tmin"], d["Tmax"], R)
        sweeps = args.ref_sweeps

        def step(n):
            return cpu_pt_baseline(md, T_all, n, threads)
        one, warm = sweeps, 10
        sample = (f"reference parallel-tempering loop, {R} temperatures on {threads} host threads, {sweeps} sweeps per "
                  f"step (swap 50, OR 10, total_energy after every Metropolis sweep, configurations swapped); "
                  f"C restatement, not Julia")
        cfg.update(replicas=R, swap_rate=50, overrelaxation_rate=10, sweeps_per_step=sweeps, l2="n/a (host)")
    else:
        def step(n):
            return cpu_baseline(md, 1.0, args.or_per_cycle, args.metro_per_cycle, n, threads)
        one, warm = args.ref_cycles, 1
        sample = (f"{threads} independent replicas (one per host thread, as one MPI rank per temperature) x "
                  f"{one} cycle(s) of the {cfg['workload']} lattice per step; reference algorithm unchanged "
                  f"(C restatement, not Julia)")
        cfg.update(cycle=f"{args.or_per_cycle} OR + {args.metro_per_cycle} Metropolis", l2="n/a (host)")
    for _ in range(args.warmup):          # untimed: a short pass (pages in the lattice and the thread pool)
        step(warm)
    tot_u, tot_t = 0.0, 0.0
    for _ in range(args.steps):
        v, u, dt = step(one)
        tot_u += u
        tot_t += dt
    value = tot_u / max(tot_t, 1e-30)
    line = {"impl": "reference", "metric": "single-spin updates/sec (Metropolis+overrelax)" + (", parallel tempering" if pt else ""),
            "value": value,
            "unit": "updates/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * tot_t / max(args.steps, 1), "higher_is_better": True, "scaling": "strong" if pt else "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
            "cpu_baseline": {"value": value, "unit": "updates/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


PT_DEFAULTS = {"C3": dict(R=64, Tmin=0.01, Tmax=1.0), "C4": dict(R=128, Tmin=0.09 / 11.6, Tmax=14 / 11.6)}


def run_pt(workload, L, world, rank, local_rank, stream, steps, warmup, sweeps_per_step, R_total=None):
    """Parallel tempering (BASELINE configs[2]/[3]): R temperature slots block-partitioned over the
    ranks' GPUs, per-replica energies gathered with NCCL on the sweep stream, temperatures exchanged
    (csmc_pt_run).  swap_rate=50, overrelaxation_rate=10 (examples/parallel_tempering/input_file.jl:19-25)."""
    import torch
    import torch.distributed as dist

    from classicalspinmc.jl_b200 import _lib, parallel
    md, cfg = workload_model(workload, L)
    d = PT_DEFAULTS[workload]
    R_total = R_total or d["R"]
    if R_total % world:
        raise SystemExit("replica count must divide over the GPUs")
    R = R_total // world
    T_all = np.geomspace(d["Tmin"], d["Tmax"], R_total)
    eng = _lib.Engine(md, n_replicas=R, seed=12345, device=local_rank, stream=stream.cuda_stream, replica_base=rank * R)
    eng.randomize(999)
    if world > 1:
        eng.comm_init(world, rank, parallel.broadcast_unique_id(_lib.comm_unique_id))
    eng.pt_init(T_all)
    p = dict(t_thermalization=10 ** 9, t_measurement=0, probe_rate=2000, swap_rate=50, overrelaxation_rate=10)
    sweep = 0
    for _ in range(max(warmup, 1)):
        eng.pt_run(p, sweep, sweep + sweeps_per_step)
        sweep += sweeps_per_step
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    l0 = eng.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        eng.pt_run(p, sweep, sweep + sweeps_per_step)
        sweep += sweeps_per_step
    e1.record(stream)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    n_metro = steps * sweeps_per_step // 10
    updates = (steps * sweeps_per_step + n_metro) * float(eng.N) * R_total
    _, ex = eng.pt_stats()
    n_col = eng.n_colours
    peak, kind = measured_peak_gbs()
    value = updates / (ms * 1e-3)
    cfg.update(replicas=R_total, replicas_per_gpu=R, swap_rate=50, overrelaxation_rate=10, sweeps_per_step=sweeps_per_step,
               colours=n_col, kernel_mode=eng.kernel_mode, sweep_groups=eng.sweep_groups()[0],
               replica_blocks=eng.replica_blocks()[0],
               exchange="temperatures swapped; "
                        + {0: "single GPU", 1: "energies via ncclAllGather", 2: "energies via peer-memory stores (push + wait kernels)",
                           3: "energies via peer-memory stores from the energy reduction kernel"}[eng.comm_mode()])
    return {"metric": "single-spin updates/sec (Metropolis+overrelax), parallel tempering", "value": value, "unit": "updates/s",
            "n_gpus": world, "steps": steps, "ms_per_step": ms / steps, "config": cfg,
            "exchanges_accepted": float(ex.sum()), "gpu_launches": int(eng.launches - l0),
            "roofline_frac_of_updates": value / world * B_ALG.get(n_col, 24.0 * (n_col + 1)) / (peak * 1e9)}


def run_ours(args):
    import torch
    import torch.distributed as dist

    from classicalspinmc.jl_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    if args.workload in PT_DEFAULTS:
        stream = torch.cuda.Stream()
        torch.cuda.set_stream(stream)
        with ClockSampler(local_rank) as clk:
            line = run_pt(args.workload, args.L, world, rank, local_rank, stream, args.steps, args.warmup, args.sweeps_per_step, args.replicas)
        if rank == 0:
            line.update(warmup=max(args.warmup, 1), higher_is_better=True, scaling="strong", vs_baseline=None, dtype="f64",
                        data="synthetic", clocks=clk.summary())
            print(json.dumps(line))
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return line

    md, cfg = workload_model(args.workload, args.L)
    # a real (non-NULL) stream: a NULL cudaStream_t in csmc_opts means "library creates its own", and
    # torch.cuda.Event only sees work on the stream it is recorded on
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    eng = _lib.Engine(md, n_replicas=1, seed=12345 + rank, device=local_rank, stream=stream.cuda_stream)
    N = eng.N
    n_col = eng.n_colours
    orc_, mc_ = args.or_per_cycle, args.metro_per_cycle
    updates_per_step = args.cycles_per_step * (orc_ + mc_) * N
    eng.randomize(12345 + rank)
    eng.set_temperatures(1.0)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput ("value") --------------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        eng.cycles_async(args.cycles_per_step, orc_, mc_)
    barrier()
    launches0 = eng.launches
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    with ClockSampler(local_rank) as clk:
        barrier()
        for k in range(args.steps):
            flush.zero_()                       # evict the lattice from L2 between timed steps
            ev[k][0].record(stream)
            eng.cycles_async(args.cycles_per_step, orc_, mc_)
            ev[k][1].record(stream)
        barrier()
    gpu_launches = eng.launches - launches0
    ms = sum(a.elapsed_time(b) for a, b in ev)
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = world * args.steps * updates_per_step / (ms * 1e-3)

    # ---- dominant kernel: the overrelaxation colour pass, timed live ------------------------------------
    n_or = 200
    eng.cycles_async(20, 1, 0)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    flush.zero_()
    e0.record(stream)
    eng.cycles_async(n_or, 1, 0)
    e1.record(stream)
    barrier()
    pass_ms = e0.elapsed_time(e1) / (n_or * n_col)
    bytes_per_launch = B_ALG.get(n_col, 24.0 * (n_col + 1)) * N / n_col
    peak, peak_kind = measured_peak_gbs()
    achieved = bytes_per_launch / (pass_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": "csmc_sweep_c<colour>_u0 (overrelaxation colour pass)", "achieved": achieved, "peak": peak,
                "peak_kind": peak_kind, "unit": "GB/s", "frac": achieved / peak,
                "us_per_launch": pass_ms * 1e3, "algorithmic_bytes_per_launch": bytes_per_launch,
                "traffic": ncu_traffic(cfg["workload"], cfg.get("L"))}

    # ---- end to end through the C-ABI with HOST buffers ---------------------------------------------------
    host_in = torch.empty((N, 3), dtype=torch.float64).pin_memory()
    host_out = torch.empty((N, 3), dtype=torch.float64).pin_memory()
    host_in.copy_(torch.from_numpy(eng.get_spins()))
    hin, hout = host_in.numpy(), host_out.numpy()

    def e2e_step():
        eng.set_spins(hin)                                  # H2D of the step's input configuration
        eng.cycles_async(args.cycles_per_step, orc_, mc_)
        eng.get_spins(out=hout)                             # D2H of the resulting configuration
        return eng.total_energy()[0]                        # + the step's scalar result

    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    e2e_value = world * args.steps * updates_per_step / float(tt.item())

    line = None
    if rank == 0:
        sk_usable, sk_rows, _, sk_budget = eng.skew_info()
        cfg.update(cycle=f"{orc_} OR + {mc_} Metropolis", cycles_per_step=args.cycles_per_step,
                   time_skewed_strips=bool(sk_usable and sk_budget < sk_rows),
                   colours=n_col, kernel_mode=eng.kernel_mode, parallelism=f"replicas x{world}",
                   launch_autotune=dict(zip(("ms_plain", "ms_pdl", "pdl_selected"), eng.autotune_report())),
                   l2=f"flushed between timed steps (256 MiB memset); lattice is {N * 24 / 2 ** 20:.0f} MiB "
                      + ("(L2-resident within a step)" if N * 24 < 100 * 2 ** 20 else
                         "(larger than L2: the colour pass is HBM-bound; sweep sequences run strip by strip through L2)"
                         if sk_usable and sk_budget < sk_rows else "(larger than L2: HBM-bound)"))
        line = {"metric": "single-spin updates/sec (Metropolis+overrelax)", "value": value, "unit": "updates/s",
                "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
                "roofline": roofline, "clocks": clk.summary(), "gpu_launches": int(gpu_launches),
                "e2e": {"value": e2e_value, "unit": "updates/s", "h2d_bytes_per_step": int(N * 24),
                        "d2h_bytes_per_step": int(N * 24 + 8)},
                "roofline_frac_of_updates": value / world * B_ALG.get(n_col, 24.0 * (n_col + 1)) / (peak * 1e9)}
        if world == 1 and not args.no_cpu_baseline:
            threads = args.cpu_threads or (os.cpu_count() or 1)
            n_ref = max(4 * args.ref_cycles, 16)          # ~10-20 s of CPU work
            v, u, dtc = cpu_baseline(md, 1.0, orc_, mc_, n_ref, threads)
            line["cpu_baseline"] = {"value": v, "unit": "updates/s", "cores": threads, "kind": "port",
                                    "sample": f"{threads} independent replicas x {n_ref} cycle(s) of the same "
                                              f"lattice ({u:.3g} updates, {dtc:.1f} s); C restatement of the reference algorithm"}
    if world > 1 and not args.no_pt:
        # the path that actually exchanges data between GPUs: parallel tempering (configs[2]), short run
        eng.close()
        pt = run_pt("C3", None, world, rank, local_rank, stream, 3, 1, 550)
        if rank == 0:
            line["pt"] = pt
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="C2")
    ap.add_argument("--L", type=int, default=None)
    ap.add_argument("--cycles-per-step", type=int, default=100,
                    help="cycles per step: 100 x (10 OR + 1 Metropolis) = 1100 sweeps, ~1 % of one annealing temperature of the README example")
    ap.add_argument("--or-per-cycle", type=int, default=10)
    ap.add_argument("--metro-per-cycle", type=int, default=1)
    ap.add_argument("--ref-cycles", type=int, default=4, help="cycles per host thread per step in the CPU legs")
    ap.add_argument("--ref-sweeps", type=int, default=55, help="PT workloads: sweeps per step in the reference arm")
    ap.add_argument("--cpu-threads", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-pt", action="store_true", help="N > 1: skip the extra parallel-tempering measurement")
    ap.add_argument("--sweeps-per-step", type=int, default=550, help="PT workloads: sweeps per timed step")
    ap.add_argument("--replicas", type=int, default=None, help="PT workloads: total replicas")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
