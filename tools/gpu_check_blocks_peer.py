#!/usr/bin/env python
"""One short single-GPU check of two launch-sequencing features (no torch import: a fresh box pages torch in
for a minute): (1) replica blocks (csmc_replica_blocks) leave results bit-identical and what they do to the
throughput of the HBM-bound PT workloads; (2) the peer-memory gather kernels (CSMC_PEER_GATHER) on a
single-rank communicator reproduce the NCCL path.  Prints one JSON line per item."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from classicalspinmc.jl_b200 import _lib  # noqa: E402
from classicalspinmc.jl_b200._abi import FLAG_JIT, FLAG_NO_RESIDENT, ModelData  # noqa: E402
from classicalspinmc.jl_b200 import workloads as models  # noqa: E402


def out(**kw):
    print(json.dumps(kw), flush=True)


def setenv(**kw):
    for k, v in kw.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = str(v)


def small_run(blocks, groups, flags):
    setenv(CSMC_REPLICA_BLOCKS=blocks, CSMC_SWEEP_GROUPS=groups)
    md = ModelData(models.kitaev_honeycomb(J3=0.25), (16, 12), 1.0)
    R = 7
    T = np.geomspace(0.2, 2.0, R)
    p = dict(t_thermalization=60, t_measurement=120, probe_rate=10, swap_rate=5, overrelaxation_rate=5)
    eng = _lib.Engine(md, n_replicas=R, seed=31, flags=flags)
    nb = eng.replica_blocks()[0]
    eng.randomize(3)
    eng.set_temperatures(T)
    eng.cycles_async(4, 5, 1)
    eng.cycles_async(1, 0, 3)
    eng.sync()
    acc = eng.accepted().copy()
    eng.pt_init(T)
    eng.pt_run(p, 0, 180)
    E, M = eng.pt_series()
    res = (np.stack([eng.get_spins(r) for r in range(R)]), acc, E, M, eng.pt_slots())
    eng.close()
    return nb, res


def check_blocks_identical():
    flags = FLAG_JIT | FLAG_NO_RESIDENT
    base = small_run(1, 1, flags)
    ok = True
    for blocks, groups in ((2, 1), (3, 2), (7, 1)):
        nb, res = small_run(blocks, groups, flags)
        same = all(np.array_equal(a, b) for a, b in zip(res, base[1]))
        ok = ok and same and nb == blocks
        out(check="replica_blocks_identical", blocks=blocks, groups=groups, in_use=nb, identical=bool(same))
    setenv(CSMC_REPLICA_BLOCKS=None, CSMC_SWEEP_GROUPS=None)
    return ok


def time_workload(name, builder, shape, S, R, blocks, n_cycles=20):
    setenv(CSMC_REPLICA_BLOCKS=blocks)
    md = ModelData(builder(), shape, S)
    t0 = time.perf_counter()
    eng = _lib.Engine(md, n_replicas=R, seed=1)
    t_create = time.perf_counter() - t0
    eng.randomize(7)
    eng.set_temperatures(np.geomspace(0.5, 2.0, R))
    eng.cycles_async(3, 10, 1)
    eng.sync()
    best = 1e30
    for _ in range(3):
        t0 = time.perf_counter()
        eng.cycles_async(n_cycles, 10, 1)
        eng.sync()
        best = min(best, time.perf_counter() - t0)
    upd = n_cycles * 11.0 * eng.N * R
    nb, ms = eng.replica_blocks()
    out(check="replica_blocks_timing", workload=name, R=R, forced=blocks, blocks_in_use=nb, autotune_ms=ms,
        groups=eng.sweep_groups()[0], Gupd_s=upd / best / 1e9, create_s=round(t_create, 2), spins_MiB=round(eng.N * R * 24 / 2 ** 20, 1))
    eng.close()
    setenv(CSMC_REPLICA_BLOCKS=None)


def pt_single_rank(peer):
    setenv(CSMC_PEER_GATHER=peer if peer else None)
    md = ModelData(models.kitaev_honeycomb(), (16, 16), 1.0)
    R = 6
    T = np.geomspace(0.1, 1.5, R)
    p = dict(t_thermalization=100, t_measurement=200, probe_rate=10, swap_rate=5, overrelaxation_rate=5)
    eng = _lib.Engine(md, n_replicas=R, seed=99, flags=FLAG_JIT | FLAG_NO_RESIDENT)
    eng.randomize(11)
    eng.comm_init(1, 0, _lib.comm_unique_id())
    mode = eng.comm_mode()
    eng.pt_init(T)
    eng.pt_run(p, 0, 150)
    eng.pt_run(p, 150, 300)
    E, M = eng.pt_series()
    acc, ex = eng.pt_stats()
    res = (E, M, eng.pt_slots(), acc, ex, np.stack([eng.get_spins(r) for r in range(R)]))
    eng.close()
    setenv(CSMC_PEER_GATHER=None)
    return mode, res


def check_peer():
    try:
        m0, base = pt_single_rank(0)
    except Exception as e:   # NCCL not loadable on this box, ...
        out(check="peer_gather", error=str(e)[:300])
        return False
    ok = True
    for peer in (1, 2):
        try:
            m, res = pt_single_rank(peer)
            same = all(np.array_equal(a, b) for a, b in zip(res, base))
            out(check="peer_gather", requested=peer, comm_mode=m, baseline_mode=m0, identical=bool(same), exchanges=float(base[4].sum()))
            ok = ok and same and m == 1 + peer
        except Exception as e:
            out(check="peer_gather", requested=peer, error=str(e)[:300])
            ok = False
    return ok


def main():
    t0 = time.perf_counter()
    what = sys.argv[1:] or ["identical", "timing-C4", "peer", "timing-C3"]
    cases = {"timing-C4": [("C4", models.pyrochlore_local, (32, 32, 32), 0.5, 128)],
             "timing-C3": [("C3", models.kitaev_honeycomb, (256, 256), 1.0, 64), ("C3", models.kitaev_honeycomb, (256, 256), 1.0, 32)]}
    for item in what:
        if item == "identical":
            out(check="replica_blocks_identical_all", ok=bool(check_blocks_identical()), t=round(time.perf_counter() - t0, 1))
        elif item == "peer":
            out(check="peer_gather_all", ok=bool(check_peer()), t=round(time.perf_counter() - t0, 1))
        elif item.startswith("sweep-"):       # throughput against the number of replica blocks
            for name, builder, shape, S, R in cases["timing-" + item[6:]]:
                for blocks in (1, 2, 3, 4, 6, 8, 12, 16, 32):
                    if blocks <= R:
                        time_workload(name, builder, shape, S, R, blocks)
        else:
            for name, builder, shape, S, R in cases[item]:
                for blocks in (1, None):
                    time_workload(name, builder, shape, S, R, blocks)
            out(t=round(time.perf_counter() - t0, 1))

if __name__ == "__main__":
    main()
