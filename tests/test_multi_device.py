"""One process, several GPUs, through the C ABI (b200_init_multi): needs >= 2 devices, skipped otherwise."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _device_count():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.gpu
@pytest.mark.parametrize("nd", [2, 4, 8])
def test_one_process_many_devices_vs_oracle(nd):
    if _device_count() < nd:
        pytest.skip("needs %d GPUs" % nd)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "multi_device_check.py"), str(nd)], capture_output=True, text=True, timeout=1200)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "multi-device check passed" in r.stdout


@pytest.mark.gpu
def test_cpp_two_devices_from_one_process():
    """tests/cpp/test_multi.cpp: a plain C++ caller (no Python, no torch) drives two devices from one process."""
    if _device_count() < 2:
        pytest.skip("needs 2 GPUs")
    exe = os.path.join(ROOT, "tests", "cpp", "test_multi")
    assert os.path.exists(exe), "run __graft_entry__.build() first"
    r = subprocess.run([exe, "2"], capture_output=True, text=True, timeout=600, env=dict(os.environ, B200_SHARD_MIN_LOGN="14"))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
