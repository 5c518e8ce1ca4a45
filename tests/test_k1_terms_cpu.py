"""K1's element-term table (csrc/srbd_model.cuh: babt_term) against the dense definition of BAbt, on the CPU:
tests/cpu_progs/check_babt_terms.cu is compiled by nvcc as a HOST program (the function is __host__ __device__)."""
import os
import shutil
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def test_babt_terms_match_dense_definition(tmp_path):
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    if not os.path.exists(nvcc):
        nvcc = shutil.which("nvcc")
    if not nvcc:
        pytest.skip("nvcc not available")
    exe = str(tmp_path / "check_babt_terms")
    src = os.path.join(HERE, "cpu_progs", "check_babt_terms.cu")
    r = subprocess.run([nvcc, "-std=c++17", "--expt-relaxed-constexpr", "-O1", "-gencode",
                        "arch=compute_100a,code=sm_100a", src, "-o", exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and "OK" in r.stdout, r.stdout + r.stderr
