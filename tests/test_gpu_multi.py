"""Multi-GPU parallel tempering (NCCL over NVLink): needs >= 2 GPUs on the box, otherwise skipped.
Launches tests/pt_worker.py under torchrun, one process per GPU."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.parametrize("world,split", [(2, "even"), (2, "uneven"), (4, "even"), (8, "even"), (8, "uneven")])
def test_sharded_parallel_tempering_is_bit_identical_to_single_gpu(world, split):
    if _n_gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29500 + world + (20 if split == "uneven" else 0)),
           os.path.join(ROOT, "tests", "pt_worker.py"), split]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    line = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert line["ok"] and line["world"] == world and line["exchanges"] > 0


@pytest.mark.parametrize("world,split,kernels,peer", [(2, "even", "resident", 1), (2, "uneven", "passes", 1), (2, "even", "passes", 2),
                                                      (8, "even", "passes", 2), (8, "uneven", "resident", 1)])
def test_peer_memory_gather_is_bit_identical_to_single_gpu(world, split, kernels, peer):
    """CSMC_PEER_GATHER=1 / 2: the measurement records travel by stores into CUDA-IPC-mapped peer memory instead of
    ncclAllGather (csmc_comm_mode 2 / 3).  Run on a B200 pair in round 1 (profiles/r1c_peer_2gpu.log) and on 8 B200s in
    round 2 (profiles/r2j_pytest_multi_8gpu.log)."""
    if _n_gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29540 + world + peer + (20 if split == "uneven" else 0)),
           os.path.join(ROOT, "tests", "pt_worker.py"), split, kernels]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, CSMC_PEER_GATHER=str(peer)))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    line = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert line["ok"] and line["world"] == world and line["exchanges"] > 0
    assert line["comm_mode"] == 1 + peer, "peers could not be mapped: the job fell back to the NCCL collectives"


def test_parallel_tempering_driver_across_processes(tmp_path):
    """examples/parallel_tempering/runner.jl through the host mirror on 2 GPUs: slots block-partitioned over
    processes, files per temperature slot written by whichever process holds the slot's replica."""
    if _n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29521", os.path.join(ROOT, "tests", "pt_driver_worker.py"),
           str(tmp_path) + "/"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    line = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert line["ok"] and line["E_cold"] < line["E_hot"]
