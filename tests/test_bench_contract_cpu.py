"""bench.py contract checks that need no GPU: the reference arm runs here (it times the unmodified reference binary,
or the C port when oracle/_ref was not built) and prints the JSON line the driver expects; our own arm refuses to
run without a B200 instead of falling back to the CPU."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT


def run_bench(*args, timeout=300):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + [str(a) for a in args],
                          stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=timeout, cwd=ROOT)


def test_reference_arm_json_line():
    p = run_bench("--impl", "reference", "--gpus", 1, "--steps", 1, "--warmup", 0, "--scale", 0.001)
    assert p.returncode == 0, p.stderr[-2000:]
    d = json.loads(p.stdout.strip().splitlines()[-1])
    assert d["impl"] == "reference" and d["metric"] == "input_kmers_per_sec_k51" and d["unit"] == "k-mers/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 1 and d["warmup"] == 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "k-mers/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=60, env=env, cwd=ROOT)
    assert p.returncode == 0 and p.stdout.strip() == ""


def test_own_arm_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    p = run_bench("--steps", 1, "--warmup", 0, "--scale", 0.001)
    assert p.returncode != 0
    assert "no CPU fallback" in p.stdout
