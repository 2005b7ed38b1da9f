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
