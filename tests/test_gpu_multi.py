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
    assert r.stdout.count("True") == 3
