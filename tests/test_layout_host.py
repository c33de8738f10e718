"""Host-side check of the kernel's memory layouts (no GPU): compiles tests/native/layout_check.cu with nvcc as a host
program and runs it.  The kernel forms every shared-memory / workspace pointer from these offsets, the host sizes
the launch from them: arrays must be disjoint, inside the allocation, and the fixed part must be a prefix."""
import os
import shutil
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.skipif(shutil.which("nvcc") is None, reason="nvcc not on PATH")
def test_layouts_are_disjoint_and_sized(tmp_path):
    exe = str(tmp_path / "layout_check")
    subprocess.check_call(["nvcc", "-std=c++17", "-O1", "-gencode", "arch=compute_100a,code=sm_100a", "-fmad=false", "-o", exe, os.path.join(HERE, "native", "layout_check.cu")])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "0 problems" in out.stdout
