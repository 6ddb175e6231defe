"""Multi-GPU parity: launches tests/mp_worker.py under torch.distributed.run.

With >= 8 (>= 4, >= 2) GPUs the ranks get one GPU each (NCCL bootstrap, NVLink peer memory).  On a 1-GPU box the SAME distributed code
(peer layer, fused depth exchange, pushes, flags) runs with all ranks sharing cuda:0 -- CUDA IPC works between processes on one
device, the GPU time-slices between them -- so the multi-rank schedules are exercised wherever the GPU tests run."""
import os, subprocess, sys
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(nproc, same_device=False, timeout=900):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
           "--master-port", str(29541 + nproc), os.path.join(ROOT, "tests", "mp_worker.py")]
    env = dict(os.environ)
    if same_device:
        env["CAPITAL_MP_SAME_DEVICE"] = "1"
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env)
    assert r.returncode == 0 and "MP_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


def test_cholinv_2x2x2_and_cacqr_on_8_gpus():
    if torch.cuda.device_count() < 8:
        pytest.skip("needs 8 GPUs")
    _run(8)


def test_cholinv_2x1x1_and_cacqr_on_2_gpus():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    _run(2)


def test_cholinv_1x2x2_on_4_gpus():
    if torch.cuda.device_count() < 4:
        pytest.skip("needs 4 GPUs")
    _run(4)


@pytest.mark.parametrize("nproc", [2, 4, 8])
def test_distributed_schedules_with_ranks_sharing_one_gpu(nproc):
    """2x1x1 (n split + final-tile broadcast), 1x2x2 (mirrors, pushes, two k classes) and the reference's 2x2x2 grid (k split,
    partial exchange fused into the GEMM epilogue) with every rank on cuda:0: golden dumps of the reference, the oracle, the
    host-pointer path, SUMMA, 1D and 3D CholeskyQR2."""
    _run(nproc, same_device=True, timeout=1500)
