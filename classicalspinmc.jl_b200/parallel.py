"""Process-group plumbing for parallel tempering: one process per GPU over ``torch.distributed``
(the stand-in for the reference's MPI.jl layer, src/monte_carlo.jl:85-92,246-256).

Only host-side bookkeeping lives here — which temperature slots a rank owns, the NCCL unique-id
bootstrap, and how per-slot results are attributed — so it is testable with the ``gloo`` backend on
CPU.  The data path (per-replica energies) is gathered on the device by libcsmc through NCCL.
"""
from __future__ import annotations

import numpy as np


def _dist():
    try:
        import torch.distributed as dist
    except Exception:       # torch is plumbing only; single-process use does not need it
        return None
    return dist if dist.is_available() and dist.is_initialized() else None


def comm_info():
    """(rank, world_size) — the analogue of MPI.Comm_rank / Comm_size (src/monte_carlo.jl:85-92)."""
    d = _dist()
    return (d.get_rank(), d.get_world_size()) if d else (0, 1)


def gather_temperatures(local_T):
    """MPI.Allgather! of the temperatures (src/monte_carlo.jl:252-254), generalised to several
    temperatures per rank.  Returns (T_all, replica_base, counts)."""
    local_T = [float(t) for t in np.atleast_1d(local_T)]
    d = _dist()
    if d is None:
        return np.array(local_T), 0, [len(local_T)]
    parts = [None] * d.get_world_size()
    d.all_gather_object(parts, local_T)
    counts = [len(p) for p in parts]
    base = int(sum(counts[: d.get_rank()]))
    return np.array([t for p in parts for t in p], dtype=np.float64), base, counts


def broadcast_unique_id(make_id):
    """Rank 0 creates the NCCL unique id (``make_id()`` -> 128 bytes), everyone receives it."""
    d = _dist()
    if d is None:
        return make_id()
    box = [make_id() if d.get_rank() == 0 else None]
    d.broadcast_object_list(box, src=0)
    return box[0]


def barrier():
    d = _dist()
    if d is not None:
        d.barrier()


def allgather_objects(obj):
    d = _dist()
    if d is None:
        return [obj]
    out = [None] * d.get_world_size()
    d.all_gather_object(out, obj)
    return out


def pairing(n_slots: int, k: int):
    """Slot pairs attempted at exchange step k = sweep / swap_rate (src/monte_carlo.jl:311-317):
    k even: (0,1),(2,3),...; k odd: (1,2),(3,4),...; unpaired ends skip."""
    first = 0 if k % 2 == 0 else 1
    return [(a, a + 1) for a in range(first, n_slots - 1, 2)]


def apply_exchanges(slot_of_replica, accepted_first_slots):
    """Host mirror of the slot permutation the device applies on accepted exchanges: the replicas in
    slots (a, a+1) trade slots for every a in ``accepted_first_slots``."""
    slot_of_replica = np.array(slot_of_replica, dtype=np.int64)
    rep_of_slot = np.argsort(slot_of_replica)
    for a in accepted_first_slots:
        ra, rb = rep_of_slot[a], rep_of_slot[a + 1]
        slot_of_replica[ra], slot_of_replica[rb] = a + 1, a
        rep_of_slot[a], rep_of_slot[a + 1] = rb, ra
    return slot_of_replica


def owner_of_replica(replica: int, counts):
    """Rank that holds global replica ``replica`` under the block partition ``counts``."""
    edges = np.cumsum([0] + list(counts))
    return int(np.searchsorted(edges, replica, side="right") - 1)


def collect_by_slot(local_items, local_slots, n_slots):
    """All ranks contribute {slot: item} for the replicas they hold; every rank receives the list
    ordered by slot (used for the final per-temperature configurations)."""
    merged = {}
    for part in allgather_objects(dict(zip([int(s) for s in local_slots], local_items))):
        merged.update(part)
    return [merged.get(s) for s in range(n_slots)]


def pt_selfcheck(world: int, rank: int, device: int, split: str = "even", flags: int | None = None):
    """Runs a small parallel-tempering job (honeycomb Kitaev-Gamma 8x8, 4 (or 3/5 alternating, ``split="uneven"``)
    temperature slots per rank, 800 sweeps, swap every 10) sharded over the ranks of the initialised process group and,
    on rank 0, the same job with every replica in one handle; the series, slot permutation, acceptance / exchange counts
    and final spins must agree bit for bit (same global replica ids -> same Philox streams and exchange decisions,
    src/monte_carlo.jl:308-349).  Every rank returns the same dict {"ok", "world", "exchanges", "comm_mode", ...}.
    Used by bench.py at N > 1 ("pt_bit_identical") and by tests/test_gpu_multi.py."""
    from . import _abi, _lib, workloads
    from ._abi import ModelData
    md = ModelData(workloads.kitaev_honeycomb(), (8, 8), 1.0)
    per_rank = [3 + 2 * (g % 2) for g in range(world)] if split == "uneven" else [4] * world
    R_total, R, base = sum(per_rank), per_rank[rank], sum(per_rank[:rank])
    T_all = np.geomspace(0.1, 1.5, R_total)
    p = dict(t_thermalization=200, t_measurement=600, probe_rate=20, swap_rate=10, overrelaxation_rate=5)
    seed = 2718
    if flags is None:
        flags = _abi.FLAG_JIT | _abi.FLAG_NO_RESIDENT      # the per-colour pass kernels: the energy reduction feeds the gather
    eng = _lib.Engine(md, n_replicas=R, seed=seed, device=device, replica_base=base, flags=flags)
    eng.randomize(500)                                     # Philox stream keyed by the global replica id
    if world > 1:
        eng.comm_init(world, rank, broadcast_unique_id(_lib.comm_unique_id))
    eng.pt_init(T_all)
    eng.pt_run(p, 0, 400)
    eng.pt_run(p, 400, 800)
    E, M = eng.pt_series()
    slots = eng.pt_slots()
    acc, ex = eng.pt_stats()
    spins = [eng.get_spins(r) for r in range(R)]
    comm_mode, kernel_mode = eng.comm_mode(), eng.kernel_mode
    eng.close()
    gathered = allgather_objects((E.tolist(), M.tolist(), slots.tolist(), acc.tolist(), ex.tolist()))
    agree = all(g == gathered[0] for g in gathered)
    all_spins = allgather_objects(spins)
    res = None
    if rank == 0:
        ref = _lib.Engine(md, n_replicas=R_total, seed=seed, device=device, replica_base=0, flags=flags)
        ref.randomize(500)
        ref.pt_init(T_all)
        ref.pt_run(p, 0, 800)
        E1, M1 = ref.pt_series()
        a1, e1 = ref.pt_stats()
        flat = [s for part in all_spins for s in part]
        same = (np.array_equal(E1, E) and np.array_equal(M1, M) and np.array_equal(ref.pt_slots(), slots)
                and np.array_equal(a1, acc) and np.array_equal(e1, ex)
                and all(np.array_equal(ref.get_spins(r), flat[r]) for r in range(R_total)))
        ref.close()
        res = {"ok": bool(agree and same and ex.sum() > 0), "world": world, "split": split, "replicas": R_total,
               "exchanges": float(ex.sum()), "probes": int(E.shape[0]), "comm_mode": comm_mode, "kernel_mode": kernel_mode,
               "ranks_agree": bool(agree), "equals_single_gpu": bool(same)}
    d = _dist()
    if d is not None:
        box = [res]
        d.broadcast_object_list(box, src=0)
        res = box[0]
    return res
