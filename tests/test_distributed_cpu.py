"""N > 1 host path on CPU: two processes over the gloo backend exercise the parallel-tempering
plumbing (temperature gather, slot partition, NCCL-id bootstrap, redundant exchange decisions,
per-slot collection) without a GPU.  The device-side gather (NCCL) is covered by
tests/test_gpu_multi.py on the GPU box."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist

    from classicalspinmc.jl_b200 import parallel
    from oracle import oracle as orc

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        assert parallel.comm_info() == (rank, world)
        # each rank owns 3 temperature slots (block partition, as one MPI rank per temperature x3)
        T_full = np.geomspace(0.1, 2.0, 3 * world)
        T_all, base, counts = parallel.gather_temperatures(T_full[3 * rank:3 * rank + 3])
        assert np.allclose(T_all, T_full) and base == 3 * rank and counts == [3] * world
        # unequal blocks (rank g holds 2 + g slots): bases are the running sums
        per = [2 + g for g in range(world)]
        T_un = np.geomspace(0.1, 2.0, sum(per))
        lo = sum(per[:rank])
        Tu, bu, cu = parallel.gather_temperatures(T_un[lo:lo + per[rank]])
        assert np.allclose(Tu, T_un) and bu == lo and cu == per
        assert [parallel.owner_of_replica(r, cu) for r in range(sum(per))] == [g for g in range(world) for _ in range(per[g])]
        # NCCL unique-id bootstrap: rank 0 creates, everyone receives the same 128 bytes
        uid = parallel.broadcast_unique_id(lambda: bytes(range(128)))
        assert uid == bytes(range(128))
        # redundant exchange decisions: every rank gathers all energies and derives the same slot
        # permutation from the shared counter-based stream (what k_pt_exchange does on the device)
        seed, n_slots = 4242, len(T_all)
        rng = np.random.default_rng(100 + rank)
        slot_of_rep = np.arange(n_slots)
        history = []
        for k in range(6):
            E_local = rng.normal(-10, 3, 3).tolist()
            E_all = np.array([e for part in parallel.allgather_objects(E_local) for e in part])
            rep_of_slot = np.argsort(slot_of_rep)
            accepted = []
            for a, b in parallel.pairing(n_slots, k):
                r4 = orc.philox(seed, a, 0xFFFFFFFF, k, 3)
                u = float(((int(r4[0]) << 32 | int(r4[1])) >> 11) * 2.0 ** -53)
                if orc.exchange_accept(T_all[a], E_all[rep_of_slot[a]], T_all[b], E_all[rep_of_slot[b]], u):
                    accepted.append(a)
            slot_of_rep = parallel.apply_exchanges(slot_of_rep, accepted)
            history.append(slot_of_rep.tolist())
        all_hist = parallel.allgather_objects(history)
        assert all(h == all_hist[0] for h in all_hist), "ranks disagree on the slot permutation"
        assert sorted(slot_of_rep.tolist()) == list(range(n_slots))
        # per-slot collection of the final configurations
        local_slots = slot_of_rep[base:base + 3]
        items = [f"cfg-of-replica-{base + r}" for r in range(3)]
        by_slot = parallel.collect_by_slot(items, local_slots, n_slots)
        rep_of_slot = np.argsort(slot_of_rep)
        assert by_slot == [f"cfg-of-replica-{rep_of_slot[s]}" for s in range(n_slots)]
        assert parallel.owner_of_replica(int(rep_of_slot[0]), counts) in range(world)
        parallel.barrier()
        open(os.path.join(out_dir, f"ok_{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_two_process_gloo_plumbing(tmp_path):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "ok_0") and os.path.exists(tmp_path / "ok_1")


def test_bench_reference_arm_other_ranks_exit_quietly():
    """bench.py --impl reference: rank 0 prints the line, other ranks exit 0 without work."""
    import subprocess
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                        "--steps", "1", "--warmup", "0", "--L", "64"], env=env, capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.strip() == ""
    env["RANK"] = "0"
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                        "--steps", "1", "--warmup", "0", "--L", "64"], env=env, capture_output=True, text=True)
    import json
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["value"] > 0 and line["cpu_baseline"]["kind"] == "port"


def test_bench_reference_arm_contract_fields():
    """The reference arm's JSON line: same metric / unit / config keys as our arm, a cpu_baseline describing the run and
    an e2e object with zero copies; the PT workloads run the reference's parallel-tempering loop."""
    import json
    import subprocess
    for extra, metric_tail in ((["--L", "64"], "(Metropolis+overrelax)"),
                               (["--workload", "C3", "--L", "16", "--replicas", "4", "--ref-sweeps", "20"], "parallel tempering")):
        r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1"] + extra,
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-2000:]
        line = json.loads(r.stdout.strip().splitlines()[-1])
        assert line["impl"] == "reference" and line["metric"].endswith(metric_tail) and line["unit"] == "updates/s"
        assert line["steps"] == 2 and line["warmup"] == 1 and line["higher_is_better"] is True and line["dtype"] == "f64"
        assert line["value"] > 0 and line["ms_per_step"] > 0 and line["vs_baseline"] is None
        assert line["cpu_baseline"]["value"] == line["value"] and line["cpu_baseline"]["cores"] >= 1 and line["cpu_baseline"]["sample"]
        assert line["e2e"] == {"value": line["value"], "unit": "updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
        assert line["config"]["workload"] in ("C2", "C3") and line["gpu_launches"] == 0
