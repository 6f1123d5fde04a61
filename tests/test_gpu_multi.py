"""Multi-process GPU parity: N row bands on N GPUs (torch.distributed.run, NCCL for the plumbing) against the
single-process CPU oracle — not against the single-GPU run of the same library.  Skipped on boxes with one GPU."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _run(world, mode, n, k, iters, bands):
    port = 29500 + (os.getpid() % 400)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(HERE, "_multi_child.py"), mode, str(n), str(k), str(iters), bands]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    return p.stdout


@pytest.mark.parametrize("mode,n,k,iters,bands", [("p2p", 1024, 5000, 25, "equal"), ("nccl", 1024, 5000, 25, "equal"),
                                                  ("p2p", 2048, 10000, 21, "unequal")])
def test_two_gpus_row_bands_vs_oracle(mode, n, k, iters, bands):
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    _run(2, mode, n, k, iters, bands)


def test_four_gpus_row_bands_vs_oracle():
    if _ngpu() < 4:
        pytest.skip("needs 4 GPUs")
    _run(4, "p2p", 2048, 10000, 21, "unequal")


def test_eight_gpus_row_bands_vs_oracle():
    if _ngpu() < 8:
        pytest.skip("needs 8 GPUs")
    _run(8, "p2p", 4096, 20000, 21, "equal")
