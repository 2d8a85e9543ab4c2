"""The host thread pool behind the pageable-input staging of ertgpu_decode (rtlamr_b200/csrc/copy_pool.hpp), on the CPU:
random sizes and offsets through pools of 1, 2, 4 and 8 threads copy exactly the requested bytes and nothing else."""
import json
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_copy_pool_copies_exactly(tmp_path):
    gxx = shutil.which("g++")
    if not gxx:
        pytest.skip("g++ not available")
    exe = str(tmp_path / "copy_pool")
    subprocess.run([gxx, "-std=c++17", "-O2", "-pthread", "-o", exe, os.path.join(ROOT, "tests", "cpu_harness", "copy_pool.cpp")],
                   check=True, capture_output=True)
    out = json.loads(subprocess.run([exe], check=True, capture_output=True, text=True, timeout=300).stdout)
    assert out["ok"] and out["copies"] == 240
