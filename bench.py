#!/usr/bin/env python
"""bench.py — single-spin updates/sec of the sweep hot path (BASELINE.json metric).

Headline workload (config.workload = "C2"): BASELINE.json configs[1], square-lattice Heisenberg L=1024 with a
z-field, single replica per GPU, cycle = 10 overrelaxation sweeps + 1 Metropolis sweep (checkerboard, 2 colours).
A "step" is `cycles_per_step` such cycles over the lattice.  N > 1 GPUs: the path does not shard a single lattice
("replicas only"), so every rank runs an independent replica of the same workload (weak
scaling, no collective).  The path that does shard — parallel tempering, replicas block-partitioned over the GPUs,
per-replica energies gathered over NVLink, temperatures swapped — is measured in the same run at EVERY N
(including 1) for C3 and C4 and printed under "pt", together with a bit-identity check of the sharded run against
the single-GPU run ("pt_bit_identical").

    python bench.py --gpus N --steps K --warmup W                      # our arm
    python bench.py --impl reference --gpus N --steps K --warmup W     # reference algorithm on the host cores

Keys beyond the base contract:
  roofline       the dominant kernel (overrelaxation colour pass) on the headline workload, timed inside the same
                 graph shape as `value` (replays of the 10-sweep OR block).  The 24 MiB lattice is L2-resident, so
                 the bound is L2 bandwidth / latency, not HBM: `peak` is an L2-resident copy measured live.
  roofline_hbm   the same kernel family where it IS HBM-bound: C2 at L=4096 (384 MiB), pass by pass, against the
                 measured HBM copy bandwidth of MEASURED_PEAKS.json; `traffic` from the committed ncu capture.
  cpu_baseline   the reference algorithm (C restatement, performance build -O3 -march=native) on the host cores:
                 all-cores aggregate (`value`) and one thread (`single_thread`).
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
COLOURS = {"C2": 2, "C3": 2, "C4": 4, "C5": 4}
L2_NOTE = ("GPU arm: L2 flushed between timed steps (256 MiB memset), the lattice is L2-resident within a step unless "
           "it exceeds 64 MiB; CPU arm: not applicable")


_T0 = time.perf_counter()


def progress(msg):
    """stage marker on stderr (rank 0): tells where a multi-rank run is if it ever stalls"""
    if int(os.environ.get("RANK", "0")) == 0:
        print(f"[bench {time.perf_counter() - _T0:7.1f} s] {msg}", file=sys.stderr, flush=True)


def workload_model(name, L=None):
    from classicalspinmc.jl_b200 import workloads
    try:
        return workloads.workload_model(name, L)
    except ValueError as e:
        raise SystemExit(str(e))


def make_config(args):
    """The workload description both arms print (identical keys and values for the same command line)."""
    from classicalspinmc.jl_b200.workloads import PT_DEFAULTS
    md, cfg = workload_model(args.workload, args.L)
    cfg["colours"] = COLOURS[args.workload]
    if args.workload in PT_DEFAULTS:
        d = PT_DEFAULTS[args.workload]
        cfg.update(replicas=args.replicas or d["R"], T_range=[d["Tmin"], d["Tmax"]], swap_rate=50, overrelaxation_rate=10,
                   probe_rate=2000, parallelism="replicas block-partitioned over the GPUs; temperatures swapped")
    else:
        cfg.update(cycle=f"{args.or_per_cycle} OR + {args.metro_per_cycle} Metropolis",
                   parallelism="independent replicas, one per GPU (no collective)")
    cfg["l2"] = L2_NOTE
    return md, cfg


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
            self._stop.wait(0.05)

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
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(key):
    """dram bytes per launch of the dominant kernel from the committed ncu --set full capture."""
    p = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    try:
        return json.load(open(p)).get(key)
    except Exception:
        return None


# ---- CPU legs (the only places bench.py executes anything under oracle/) ---------------------------------------------
_CPU_BUILD = {"flags": None}


def cpu_lattice(md):
    """The oracle lattice of the timed CPU legs: the performance build (-O3 -march=native, compiled on this box); if no
    compiler is available here, the shipped checker build (-O3 -march=x86-64-v3 -ffp-contract=off) is timed instead and
    `cpu_baseline.sample` says so."""
    from oracle import oracle as orc
    try:
        lat = orc.OracleLattice(md, fast=True)
        _CPU_BUILD["flags"] = "gcc " + orc.FAST_CFLAGS
    except Exception as e:                                   # no gcc on the box, read-only tree, ...
        lat = orc.OracleLattice(md)
        _CPU_BUILD["flags"] = f"checker build -O3 -march=x86-64-v3 -ffp-contract=off (performance build unavailable: {type(e).__name__})"
    return lat


def cpu_cycles(md, T, or_per_cycle, metro_per_cycle, n_cycles, threads):
    """The oracle's restatement of the reference algorithm (random-site Metropolis with two energy() evaluations,
    sequential overrelaxation), performance build, timed on the host cores: `threads` independent replicas."""
    lat = cpu_lattice(md)
    spins = np.concatenate([lat.randomize(seed=12345, replica=r) for r in range(threads)])
    t0 = time.perf_counter()
    updates = lat.cycles(spins, threads, T, n_cycles, or_per_cycle, metro_per_cycle)
    dt = time.perf_counter() - t0
    return updates / dt, updates, dt


def cpu_pt(md, T_all, sweeps, threads, swap_rate=50, rate=10, seed=3):
    """The oracle's restatement of the reference parallel-tempering loop (src/monte_carlo.jl:289-349: OR every
    sweep, random-site Metropolis + total_energy every `rate`-th, configuration-swapping exchange every
    `swap_rate`-th), one temperature per host thread as examples/parallel_tempering/README.txt:17."""
    lat = cpu_lattice(md)
    R = len(T_all)
    spins = np.concatenate([lat.randomize(seed=12345, replica=r) for r in range(R)])
    t0 = time.perf_counter()
    lat.parallel_tempering(spins, T_all, sweeps, 0, 2000, swap_rate, rate, seed=seed, n_threads=threads)
    dt = time.perf_counter() - t0
    updates = float(sweeps + (sweeps + rate - 1) // rate) * lat.N * R
    return updates / dt, updates, dt


def cpu_flags():
    if _CPU_BUILD["flags"] is None:
        from oracle import oracle as orc
        return "gcc " + orc.FAST_CFLAGS
    return _CPU_BUILD["flags"]


def run_reference(args):
    """Reference arm: the reference's CPU algorithm for the workload (C restatement under oracle/, the
    reference itself is Julia and cannot run here) on all host threads; rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from classicalspinmc.jl_b200.workloads import PT_DEFAULTS
    md, cfg = make_config(args)
    threads = args.cpu_threads or (os.cpu_count() or 1)
    pt = args.workload in PT_DEFAULTS
    if pt:
        T_all = np.geomspace(cfg["T_range"][0], cfg["T_range"][1], cfg["replicas"])
        sweeps = args.ref_sweeps

        def step(n):
            return cpu_pt(md, T_all, n, threads)
        one, warm = sweeps, 10
        sample = (f"reference parallel-tempering loop, {len(T_all)} temperatures on {threads} host threads, {sweeps} sweeps per "
                  f"step (swap 50, OR 10, total_energy after every Metropolis sweep, configurations swapped); "
                  f"C restatement, not Julia; ")
    else:
        def step(n):
            return cpu_cycles(md, 1.0, args.or_per_cycle, args.metro_per_cycle, n, threads)
        one, warm = args.ref_cycles, 1
        sample = (f"{threads} independent replicas (one per host thread, as one MPI rank per temperature) x "
                  f"{one} cycle(s) of the {cfg['workload']} lattice per step; reference algorithm unchanged "
                  f"(C restatement, not Julia); ")
    for _ in range(args.warmup):          # untimed: a short pass (pages in the lattice and the thread pool)
        step(warm)
    tot_u, tot_t = 0.0, 0.0
    for _ in range(args.steps):
        v, u, dt = step(one)
        tot_u += u
        tot_t += dt
    value = tot_u / max(tot_t, 1e-30)
    sample += cpu_flags()            # known once the library has been built / loaded
    line = {"impl": "reference", "metric": "single-spin updates/sec (Metropolis+overrelax)" + (", parallel tempering" if pt else ""),
            "value": value,
            "unit": "updates/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * tot_t / max(args.steps, 1), "higher_is_better": True, "scaling": "strong" if pt else "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
            "cpu_baseline": {"value": value, "unit": "updates/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ---- parallel tempering (the sharded path) ---------------------------------------------------------------------------
def run_pt(workload, L, world, rank, local_rank, stream, steps, warmup, sweeps_per_step, R_total=None, single=False):
    """Parallel tempering (BASELINE configs[2]/[3]): R temperature slots block-partitioned over the
    ranks' GPUs, per-replica energies gathered over NVLink on the sweep stream, temperatures exchanged
    (csmc_pt_run).  swap_rate=50, overrelaxation_rate=10 (examples/parallel_tempering/input_file.jl:19-25).
    single=True: this process alone holds every replica (the N=1 point of the strong-scaling curve)."""
    import torch
    import torch.distributed as dist

    from classicalspinmc.jl_b200 import _lib, parallel
    from classicalspinmc.jl_b200.workloads import PT_DEFAULTS
    md, _ = workload_model(workload, L)
    d = PT_DEFAULTS[workload]
    R_total = R_total or d["R"]
    if single:
        world, rank = 1, 0
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
    launches_timed = int(eng.launches - l0)
    _, ex = eng.pt_stats()
    # the exchange step alone (energy reduction of every local replica, gather of the records across the GPUs, exchange
    # decisions: what runs once per swap_rate sweeps), timed through csmc_pt_exchange with a host sync per call: an upper
    # bound of the collective's share of an exchange period
    for k in range(3):
        eng.pt_exchange(k & 1)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    n_ex = 20
    t0 = time.perf_counter()
    for k in range(n_ex):
        eng.pt_exchange(k & 1)
    torch.cuda.synchronize()
    ex_us = torch.tensor([(time.perf_counter() - t0) / n_ex * 1e6], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ex_us, op=dist.ReduceOp.MAX)
    ex_us = float(ex_us.item())
    period_us = ms / steps * 1e3 / (sweeps_per_step / 50.0)
    n_metro = steps * sweeps_per_step // 10
    updates = (steps * sweeps_per_step + n_metro) * float(eng.N) * R_total
    n_col = eng.n_colours
    peak, _ = measured_peak_gbs()
    value = updates / (ms * 1e-3)
    out = {"workload": workload, "value": value, "unit": "updates/s", "n_gpus": world, "replicas": R_total, "replicas_per_gpu": R,
           "steps": steps, "sweeps_per_step": sweeps_per_step, "ms_per_step": ms / steps,
           "exchanges_accepted": float(ex.sum()), "gpu_launches": launches_timed,
           "exchange_step_us": ex_us, "exchange_period_us": period_us, "exchange_share_upper_bound": ex_us / period_us,
           "engine": {"kernel_mode": eng.kernel_mode, "sweep_groups": eng.sweep_groups()[0], "replica_blocks": eng.replica_blocks()[0],
                      "persistent_tiles": eng.persist_info()[0] if hasattr(eng, "persist_info") else 0,
                      "exchange": {0: "single GPU", 1: "energies via ncclAllGather", 2: "energies via peer-memory stores (push + wait kernels)",
                                   3: "energies via peer-memory stores from the energy reduction kernel"}[eng.comm_mode()]},
           "hbm_roofline_frac_of_updates": value / world * B_ALG.get(n_col, 24.0 * (n_col + 1)) / (peak * 1e9)}
    eng.close()
    return out


def pt_records(world, rank, local_rank, stream, args):
    """The "pt" object of the bench line: C3 and C4 at this N, each with the single-GPU point measured in the same
    process set (rank 0 alone, N > 1 only) and the strong-scaling efficiency value / (N * single-GPU value); plus the
    bit-identity check of the sharded run against the single-GPU run on a small lattice (N > 1)."""
    import torch.distributed as dist

    from classicalspinmc.jl_b200 import parallel
    out = {}
    for wl in ("C3", "C4"):
        progress(f"parallel tempering {wl} on {world} GPU(s)")
        rec = run_pt(wl, None, world, rank, local_rank, stream, args.pt_steps, 2, args.sweeps_per_step)
        if world > 1:
            if rank == 0:
                progress(f"parallel tempering {wl}, single-GPU point on rank 0")
                one = run_pt(wl, None, world, rank, local_rank, stream, args.pt_steps, 2, args.sweeps_per_step, single=True)
                rec["single_gpu_value"] = one["value"]
                rec["pt_efficiency"] = rec["value"] / (world * one["value"])
            dist.barrier()
        else:
            rec["single_gpu_value"] = rec["value"]
            rec["pt_efficiency"] = 1.0
        out[wl] = rec
    ident = None
    if world > 1:
        checks = {}
        for split in ("even", "uneven"):
            progress(f"bit-identity check of the sharded run ({split} replica blocks)")
            checks[split] = parallel.pt_selfcheck(world, rank, local_rank, split)
        ident = {"ok": all(c["ok"] for c in checks.values()), **checks}
    return out, ident


# ---- our arm ---------------------------------------------------------------------------------------------------------
def l2_copy_peak(torch, stream, n_bytes=12 << 20, reps=50):
    """GB/s (read + write) of an L2-resident device copy (src and dst together = the lattice's 24 MiB), replayed as a
    CUDA graph so that launch gaps do not count: the live denominator of the L2-resident roofline."""
    a = torch.empty(n_bytes, dtype=torch.uint8, device="cuda")
    b = torch.empty(n_bytes, dtype=torch.uint8, device="cuda")
    a.zero_()
    for _ in range(3):
        b.copy_(a)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=stream, capture_error_mode="thread_local"):   # other threads (NCCL watchdog, clock sampler) may call CUDA
        for _ in range(reps):
            b.copy_(a)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e30
    for _ in range(3):
        e0.record(stream)
        g.replay()
        e1.record(stream)
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return 2.0 * n_bytes * reps / (best * 1e-3) / 1e9


def run_ours(args):
    import torch
    import torch.distributed as dist

    from classicalspinmc.jl_b200 import _abi, _lib
    from classicalspinmc.jl_b200.workloads import PT_DEFAULTS

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    md, cfg = make_config(args)
    # a real (non-NULL) stream: a NULL cudaStream_t in csmc_opts means "library creates its own", and
    # torch.cuda.Event only sees work on the stream it is recorded on
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0

    if args.workload in PT_DEFAULTS:
        with ClockSampler(local_rank) as clk:
            rec = run_pt(args.workload, args.L, world, rank, local_rank, stream, args.steps, args.warmup, args.sweeps_per_step, args.replicas)
        if rank == 0:
            line = {"metric": "single-spin updates/sec (Metropolis+overrelax), parallel tempering", "value": rec["value"], "unit": "updates/s",
                    "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 1), "ms_per_step": rec["ms_per_step"],
                    "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                    "config": cfg, "engine": rec["engine"], "exchanges_accepted": rec["exchanges_accepted"],
                    "gpu_launches": rec["gpu_launches"], "hbm_roofline_frac_of_updates": rec["hbm_roofline_frac_of_updates"],
                    "clocks": clk.summary()}
            print(json.dumps(line))
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    eng = _lib.Engine(md, n_replicas=1, seed=12345 + rank, device=local_rank, stream=stream.cuda_stream)
    N = eng.N
    n_col = eng.n_colours
    assert n_col == cfg["colours"]
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
    progress(f"{args.workload}: engine ready, timing {args.steps} steps")
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

    # ---- dominant kernel: the overrelaxation colour pass, timed live inside the same graph shape as `value`
    # (replays of the `orc_`-sweep OR block of the cycle graph, no Metropolis sweeps) -------------------------
    def time_or_block(e, or_sweeps, reps, sync=barrier):
        """ms per replay of an `or_sweeps`-sweep OR block.  `sync` must be the local torch.cuda.synchronize when only
        one rank takes the measurement (a collective barrier there would never be matched by the other ranks)."""
        e.cycles_async(max(reps // 5, 2), or_sweeps, 0)
        sync()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        flush.zero_()
        a.record(stream)
        e.cycles_async(reps, or_sweeps, 0)
        b.record(stream)
        sync()
        return a.elapsed_time(b) / reps          # ms per block
    progress("dominant kernel inside the OR block graph")
    or_block = max(orc_, 1)
    block_ms = time_or_block(eng, or_block, 100)
    l_before = eng.launches
    eng.cycles_async(1, or_block, 0)
    launches_per_block = int(eng.launches - l_before)          # the library counts the launches of a graph replay
    barrier()
    pass_ms = block_ms / launches_per_block
    balg = B_ALG.get(n_col, 24.0 * (n_col + 1))
    bytes_per_launch = balg * N * or_block / launches_per_block
    achieved = bytes_per_launch / (pass_ms * 1e-3) / 1e9
    hbm_peak, hbm_kind = measured_peak_gbs()
    l2_resident = N * 24 <= 64 * 2 ** 20
    if l2_resident:
        l2_peak = l2_copy_peak(torch, stream)
        roofline = {"bound": "l2", "bound_note": "the 24 MiB lattice is L2-resident within a step: the pass is bound by L2 bandwidth and "
                    "launch / load latency, not by HBM (see roofline_hbm for the HBM-bound point of the same kernel)",
                    "achieved": achieved, "peak": l2_peak, "peak_kind": "measured live: L2-resident device copy (12 MiB -> 12 MiB, graph replay), read + write bytes",
                    "unit": "GB/s", "frac": achieved / l2_peak, "frac_of_hbm_peak": achieved / hbm_peak}
    else:
        roofline = {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "peak_kind": hbm_kind, "unit": "GB/s", "frac": achieved / hbm_peak}
    # fp64 pipe roofline of the same launch: flops counted by the code generator (fma = 2) against the nominal vector
    # fp64 rate of the part (64 DFMA per SM and clock at the maximum SM clock; no measured fp64 peak exists in
    # MEASURED_PEAKS.json)
    flops_upd, _ = eng.kernel_costs()
    sm_count = torch.cuda.get_device_properties(local_rank).multi_processor_count
    fp64_peak = sm_count * 64 * 2 * (clk.max_mhz or 1965) * 1e6 / 1e12
    fp64_tflops = flops_upd * (N * or_block / launches_per_block) / (pass_ms * 1e-3) / 1e12
    roofline.update(fp64={"flops_per_update": flops_upd, "achieved_tflops": fp64_tflops, "peak_tflops_nominal": fp64_peak,
                          "frac": fp64_tflops / fp64_peak, "peak_kind": f"nominal: {sm_count} SMs x 64 DFMA/clk x max SM clock"})
    pm = eng.persist_info() if hasattr(eng, "persist_info") else (0, 0, 0)
    roofline.update(kernel=("csmc_persist (tile-resident multi-pass kernel: one launch per OR block)" if pm[0] else
                            "csmc_sweep_c<colour>_u0 (overrelaxation colour pass)"),
                    timed_as=f"replays of the {or_block}-sweep OR block graph ({launches_per_block} launch(es) per block)",
                    us_per_launch=pass_ms * 1e3, us_per_colour_pass=block_ms * 1e3 / (or_block * n_col),
                    algorithmic_bytes_per_launch=bytes_per_launch,
                    traffic=ncu_traffic(f"{cfg['workload']}:{cfg['L']}" + (":persist" if pm[0] else "")),
                    traffic_steady_state=ncu_traffic(f"{cfg['workload']}:{cfg['L']}:steady"),
                    ncu_steady_state=ncu_traffic(f"{cfg['workload']}:{cfg['L']}:ncu_steady_state"),
                    traffic_note="`traffic`: isolated launch with caches flushed by ncu (the whole lattice is read once); "
                                 "`traffic_steady_state`: the same launch inside a running cycle with the lattice L2-resident")

    # ---- end to end through the C-ABI with HOST buffers ---------------------------------------------------
    progress("end to end with host buffers")
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

    sk_usable, sk_rows, _, sk_budget = eng.skew_info()
    engine = dict(kernel_mode=eng.kernel_mode, cycles_per_step=args.cycles_per_step,
                  time_skewed_strips=bool(sk_usable and sk_budget < sk_rows), persistent_tiles=pm[0],
                  launch_autotune=dict(zip(("ms_plain", "ms_pdl", "pdl_selected"), eng.autotune_report())),
                  lattice_MiB=N * 24 / 2 ** 20)
    eng.close()

    # ---- the same kernel family where it is HBM-bound: C2 at L=4096, pass by pass ------------------------------
    roofline_hbm = None
    if args.workload == "C2" and not args.no_hbm_point and world == 1:      # a single-GPU property: measured at N = 1 only
        progress("HBM-bound point: C2 at L=4096")
        md4, _ = workload_model("C2", 4096)
        e4 = _lib.Engine(md4, n_replicas=1, seed=1, device=local_rank, stream=stream.cuda_stream, flags=_abi.FLAG_NO_AUTOTUNE)
        e4.randomize(7)
        e4.set_temperatures(1.0)
        ms4 = time_or_block(e4, 1, 40, sync=torch.cuda.synchronize) / 2   # one sweep per graph: pass by pass (no strips), 2 launches
        b4 = 72.0 * e4.N / 2
        a4 = b4 / (ms4 * 1e-3) / 1e9
        roofline_hbm = {"bound": "hbm", "workload": "C2 at L=4096 (384 MiB of spins, 3x the L2)", "kernel": "csmc_sweep_c<colour>_u0 (overrelaxation colour pass)",
                        "achieved": a4, "peak": hbm_peak, "peak_kind": hbm_kind, "unit": "GB/s", "frac": a4 / hbm_peak, "us_per_launch": ms4 * 1e3,
                        "algorithmic_bytes_per_launch": b4, "traffic": ncu_traffic("C2:4096")}
        # whole cycle at this size (time-skewed strips keep a strip of the lattice L2-resident across the 22 passes)
        e4.cycles_async(3, orc_, mc_)
        barrier_local = torch.cuda.synchronize
        barrier_local()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        e4.cycles_async(10, orc_, mc_)
        b.record(stream)
        barrier_local()
        roofline_hbm["cycle_updates_per_s"] = 10 * (orc_ + mc_) * e4.N / (a.elapsed_time(b) * 1e-3)
        roofline_hbm["cycle_time_skewed_strips"] = bool(e4.skew_info()[0])
        e4.close()

    pt, ident = (None, None)
    if not args.no_pt:
        pt, ident = pt_records(world, rank, local_rank, stream, args)

    if rank == 0:
        line = {"metric": "single-spin updates/sec (Metropolis+overrelax)", "value": value, "unit": "updates/s",
                "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg, "engine": engine,
                "roofline": roofline, "clocks": clk.summary(), "gpu_launches": int(gpu_launches),
                "e2e": {"value": e2e_value, "unit": "updates/s", "h2d_bytes_per_step": int(N * 24),
                        "d2h_bytes_per_step": int(N * 24 + 8)},
                "hbm_roofline_frac_of_updates": value / world * balg / (hbm_peak * 1e9)}
        if roofline_hbm:
            line["roofline_hbm"] = roofline_hbm
        if pt:
            line["pt"] = pt
        if ident is not None:
            line["pt_bit_identical"] = ident["ok"]
            line["pt_bit_identical_detail"] = ident
        if world == 1 and not args.no_cpu_baseline:
            progress("CPU baseline legs")
            threads = args.cpu_threads or (os.cpu_count() or 1)
            n_ref = max(2 * args.ref_cycles, 8)          # ~5-10 s of CPU work per leg
            v, u, dtc = cpu_cycles(md, 1.0, orc_, mc_, n_ref, threads)
            v1, u1, dt1 = cpu_cycles(md, 1.0, orc_, mc_, max(n_ref // 2, 2), 1)
            line["cpu_baseline"] = {"value": v, "unit": "updates/s", "cores": threads, "kind": "port",
                                    "sample": f"{threads} independent replicas x {n_ref} cycle(s) of the same "
                                              f"lattice ({u:.3g} updates, {dtc:.1f} s); C restatement of the reference algorithm; {cpu_flags()}",
                                    "single_thread": {"value": v1, "cores": 1, "sample": f"1 replica x {max(n_ref // 2, 2)} cycles ({u1:.3g} updates, {dt1:.1f} s)"}}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


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
    ap.add_argument("--no-pt", action="store_true", help="skip the parallel-tempering records (C3, C4) and the bit-identity check")
    ap.add_argument("--no-hbm-point", action="store_true", help="skip the L=4096 HBM-bound measurement of the dominant kernel")
    ap.add_argument("--pt-steps", type=int, default=5, help="timed steps of the parallel-tempering records")
    ap.add_argument("--sweeps-per-step", type=int, default=550, help="PT workloads: sweeps per timed step")
    ap.add_argument("--replicas", type=int, default=None, help="PT workloads: total replicas")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
