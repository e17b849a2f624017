"""Host-only check of the backward's operand-panel addressing for hi-only and hi+lo panels (tests/csrc/
panel_layout_check.cu compiles tc_bwd_layout.cuh's __host__ __device__ helpers with nvcc and runs them on the CPU)."""
import os
import shutil
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def test_panel_addressing_is_consistent(tmp_path):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    exe = os.path.join(tmp_path, "panel_check")
    subprocess.run([nvcc, "-std=c++17", "-Wno-deprecated-gpu-targets", "-o", exe, os.path.join(HERE, "csrc", "panel_layout_check.cu")],
                   check=True, capture_output=True)
    res = subprocess.run([exe], capture_output=True, text=True)
    assert res.returncode == 0 and "panel layout ok" in res.stdout, res.stdout
