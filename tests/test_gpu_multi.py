"""Multi-GPU parity: launches tests/mp_worker.py under torch.distributed.run when the box has >= 8 (or >= 2) GPUs."""
import os, subprocess, sys
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(nproc):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
           "--master-port", "29541", os.path.join(ROOT, "tests", "mp_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "MP_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


def test_cholinv_2x2x2_and_cacqr_1d_on_8_gpus():
    if torch.cuda.device_count() < 8:
        pytest.skip("needs 8 GPUs")
    _run(8)


def test_cacqr_1d_on_2_gpus():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    _run(2)


def test_cholinv_1x2x2_on_4_gpus():
    if torch.cuda.device_count() < 4:
        pytest.skip("needs 4 GPUs")
    _run(4)
