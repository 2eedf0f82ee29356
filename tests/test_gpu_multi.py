"""Multi-GPU parity (needs >= 2 CUDA devices; skipped otherwise): the row-sharded training forward with the exchange
fused into the finalize kernel (peer memory over NVLink), and the all-reduce route, must both leave every rank with the
codebook / cluster_size of a single-process forward on the concatenated batch (tools_dist_check.py asserts it)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("fused", ["1", "0"])
def test_two_rank_forward_matches_single_process(fused):
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    import socket
    env = dict(os.environ, MEDTOK_FUSED_EXCHANGE=fused)
    with socket.socket() as sock:                      # a free rendezvous port
        sock.bind(("127.0.0.1", 0))
        port = sock.getsockname()[1]
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tools_dist_check.py")],
                       capture_output=True, text=True, env=env, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("rank ")]
    assert len(lines) == 2
    want = "fused NVLink exchange" if fused == "1" else "all_reduce + finalize"
    assert all(want in ln and "replicas identical=True" in ln for ln in lines), lines


def test_two_rank_ddp_step_under_autocast():
    """train_MedTok.py:185,212,240-246: DDP(find_unused_parameters=True) + autocast + GradScaler around the soft quantizer"""
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    import socket
    with socket.socket() as sock:
        sock.bind(("127.0.0.1", 0))
        port = sock.getsockname()[1]
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tools_ddp_check.py")],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("rank ")]
    assert len(lines) == 2 and all("identical across ranks=True" in ln and "grads finite=True" in ln for ln in lines), lines
