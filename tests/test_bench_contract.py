"""bench.py's JSON contract, checked on CPU through the reference arm (the GPU arm needs a B200)."""
import json
import os
import subprocess
import sys

from conftest import ROOT

KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
        "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline", "impl"}


def test_reference_arm_prints_one_contract_line(built):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--cpu-sample-mib", "2"], capture_output=True, text=True, timeout=300, check=True)
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert KEYS <= set(j)
    assert j["impl"] == "reference" and j["higher_is_better"] is True and j["vs_baseline"] is None
    assert j["unit"] == "Msamples/s" and j["value"] > 1.0
    assert j["cpu_baseline"]["kind"] == "port" and j["cpu_baseline"]["cores"] >= 1
    assert j["e2e"]["value"] == j["value"] and j["e2e"]["h2d_bytes_per_step"] == 0
    assert "workload" in j["config"] and "model" not in j["config"]


def test_non_zero_ranks_of_the_reference_arm_do_no_work(built):
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                         capture_output=True, text=True, timeout=60, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""
