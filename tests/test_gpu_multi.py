"""Multi-GPU parity (needs >= 2 B200s on the box; skipped otherwise): N-GPU result == 1-GPU result."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_gpus_equal_one_gpu_bit_for_bit(lib):
    if lib.evplp_device_count() < 2:
        pytest.skip("needs 2 GPUs")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "scripts", "check_multigpu.py")],
                       capture_output=True, text=True, timeout=600)
    sys.stdout.write(r.stdout[-2000:])
    assert r.returncode == 0, r.stderr[-3000:]
    assert r.stdout.count("True") == 6  # 3 layers x 2 partition modes


def test_evplp_reduce_with_real_nccl_communicators(lib, tmp_path):
    """The C-ABI reduce entry point with raw ncclComm_t handles (one process, two GPUs)."""
    if lib.evplp_device_count() < 2:
        pytest.skip("needs 2 GPUs")
    exe = str(tmp_path / "nccl_reduce_test")
    src = os.path.join(ROOT, "tests", "hostsim", "nccl_reduce_test.cpp")
    libdir = os.path.join(ROOT, "evplp_b200", "lib")
    cc = subprocess.run(["/usr/bin/g++", "-O1", "-std=c++17", src, "-o", exe, "-I/usr/local/cuda/include", f"-L{libdir}", "-levplp_b200",
                         "-L/usr/local/cuda/lib64", "-lcudart", "-lnccl", f"-Wl,-rpath,{libdir}", "-Wl,-rpath,/usr/local/cuda/lib64"],
                        capture_output=True, text=True)
    if cc.returncode != 0:
        pytest.skip("cannot build the NCCL test program here: " + cc.stderr[-300:])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "OK" in r.stdout or "SKIP" in r.stdout
