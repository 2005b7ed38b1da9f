"""bench.py contract checks that need no GPU: both arms describe the workload with the same `config`, and the reference
arm (the reference's CPU algorithm, C restatement) prints the keys the driver reads."""
import argparse
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def _args(**kw):
    d = dict(workload="C2", L=None, replicas=None, or_per_cycle=10, metro_per_cycle=1)
    d.update(kw)
    return argparse.Namespace(**d)


@pytest.mark.parametrize("workload", ["C2", "C3", "C4", "C5"])
def test_both_arms_print_the_same_config(workload):
    """`config` comes from one function for both arms and holds nothing that depends on the engine or the step size."""
    _, a = bench.make_config(_args(workload=workload))
    _, b = bench.make_config(_args(workload=workload))
    assert a == b and a["workload"] == workload and a["colours"] == bench.COLOURS[workload]
    assert "l2" in a and "parallelism" in a
    for engine_key in ("kernel_mode", "cycles_per_step", "launch_autotune", "sweep_groups", "replica_blocks"):
        assert engine_key not in a
    json.dumps(a)


def test_reference_arm_line_has_the_contract_keys():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--L", "32", "--steps", "2", "--warmup", "1",
                        "--ref-cycles", "1", "--cpu-threads", "2"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert line["impl"] == "reference" and line["unit"] == "updates/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["steps"] == 2 and line["warmup"] == 1 and line["gpu_launches"] == 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] == 2 and line["cpu_baseline"]["value"] == line["value"]
    assert "march" in line["cpu_baseline"]["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    _, cfg = bench.make_config(_args(L=32))
    assert line["config"] == cfg


def test_reference_arm_other_ranks_do_no_work():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"], capture_output=True, text=True,
                       timeout=120, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""
