"""torchrun worker for tests/test_gpu_multi.py and tools: parallel tempering sharded over the ranks'
GPUs (NCCL gather of the per-replica energies) must reproduce, bit for bit, the single-GPU run that
holds every replica (same global replica ids -> same Philox streams, same exchange decisions)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist

    from classicalspinmc.jl_b200 import _abi, _lib, parallel
    from classicalspinmc.jl_b200._abi import ModelData
    from oracle import oracle as orc
    from tests import models

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    md = ModelData(models.kitaev_honeycomb(), (8, 8), 1.0)
    lat = orc.OracleLattice(md)
    # "uneven": the ranks hold different numbers of temperature slots (grouped ncclBroadcast instead of ncclAllGather)
    per_rank = [3 + 2 * (g % 2) for g in range(world)] if "uneven" in sys.argv[1:] else [4] * world
    R_total, R = sum(per_rank), per_rank[rank]
    T_all_expected = np.geomspace(0.1, 1.5, R_total)
    first = sum(per_rank[:rank])
    T_all, base, counts = parallel.gather_temperatures(T_all_expected[first:first + R])
    assert np.allclose(T_all, T_all_expected) and base == first and counts == per_rank
    p = dict(t_thermalization=200, t_measurement=600, probe_rate=20, swap_rate=10, overrelaxation_rate=5)
    seed = 2718
    # "passes": per-colour pass kernels instead of the resident kernel (the energy reduction then feeds the gather)
    flags = (_abi.FLAG_JIT | _abi.FLAG_NO_RESIDENT) if "passes" in sys.argv[1:] else 0
    eng = _lib.Engine(md, n_replicas=R, seed=seed, device=local, replica_base=base, flags=flags)
    for r in range(R):
        eng.set_spins(lat.randomize(seed=500 + base + r), replica=r)
    uid = parallel.broadcast_unique_id(_lib.comm_unique_id)
    eng.comm_init(world, rank, uid)
    comm_mode = eng.comm_mode()      # 1 NCCL collectives, 2 / 3 stores into peer memory (CSMC_PEER_GATHER=1 / 2)
    eng.pt_init(T_all)
    eng.pt_run(p, 0, 400)
    eng.pt_run(p, 400, 800)
    E, M = eng.pt_series()
    slots = eng.pt_slots()
    acc, ex = eng.pt_stats()
    spins = [eng.get_spins(r) for r in range(R)]
    # every rank holds the full series; they must agree exactly
    gathered = parallel.allgather_objects((E.tolist(), M.tolist(), slots.tolist(), acc.tolist(), ex.tolist()))
    assert all(g == gathered[0] for g in gathered), "ranks disagree"
    all_spins = parallel.allgather_objects(spins)
    ok = True
    if rank == 0:
        ref = _lib.Engine(md, n_replicas=R_total, seed=seed, device=local, replica_base=0, flags=flags)
        for r in range(R_total):
            ref.set_spins(lat.randomize(seed=500 + r), replica=r)
        ref.pt_init(T_all)
        ref.pt_run(p, 0, 800)
        E1, M1 = ref.pt_series()
        assert np.array_equal(E1, E) and np.array_equal(M1, M), "multi-GPU series differ from single-GPU"
        assert np.array_equal(ref.pt_slots(), slots)
        a1, e1 = ref.pt_stats()
        assert np.array_equal(a1, acc) and np.array_equal(e1, ex)
        flat = [s for part in all_spins for s in part]
        for r in range(R_total):
            assert np.array_equal(ref.get_spins(r), flat[r])
        assert ex.sum() > 0
        print(json.dumps({"ok": True, "world": world, "exchanges": float(ex.sum()), "probes": int(E.shape[0]),
                          "comm_mode": comm_mode, "kernel_mode": eng.kernel_mode}))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
