"""The drop-in boundary without Python in the process: tests/c_host/host_forward.cpp is a plain C++ host (CUDA runtime +
include/omg_b200.h, no torch) that runs a Linear and a LayerNorm through the C ABI on cudaMalloc'ed buffers, records them
into a launch plan and replays the plan on new data; it checks itself against a CPU loop and exits non-zero on a miss."""
import os
import shutil
import subprocess
import tempfile

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cpp_host_runs_gemm_layernorm_and_a_launch_plan_through_the_c_abi():
    from omg_b200 import _lib
    cxx = shutil.which("g++")
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    if cxx is None or not os.path.exists(os.path.join(cuda, "include", "cuda_runtime.h")):
        pytest.skip("no host C++ compiler / CUDA headers on this box")
    libdir = os.path.dirname(_lib.LIB_PATH)
    with tempfile.TemporaryDirectory() as tmp:
        exe = os.path.join(tmp, "host_forward")
        subprocess.run([cxx, "-O1", "-std=c++17", os.path.join(ROOT, "tests", "c_host", "host_forward.cpp"),
                        "-I", os.path.join(ROOT, "include"), "-I", os.path.join(cuda, "include"),
                        "-L", libdir, "-L", os.path.join(cuda, "lib64"), "-lomg_b200", "-lcudart",
                        f"-Wl,-rpath,{libdir}", f"-Wl,-rpath,{os.path.join(cuda, 'lib64')}", "-o", exe], check=True)
        r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
        print(r.stdout, r.stderr)
        assert r.returncode == 0, r.stdout + r.stderr
        assert "C host OK" in r.stdout and "plan length 2" in r.stdout
