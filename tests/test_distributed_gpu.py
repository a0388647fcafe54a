"""GPU test of the distributed prefix doubling (libsais_b200/dist.py): launched with torchrun in a
subprocess (NCCL), on one GPU always and on two when the box has them; each rank's slice of the
suffix array must equal the single-GPU result."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _run(nproc, args, port=None):
    port = port or _free_port()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc),
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tools", "dist_sa.py")] + args
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert r.returncode == 0 and lines, r.stderr[-2000:]
    return json.loads(lines[-1])


@pytest.mark.parametrize("args", [["20", "dna"], ["22", "bytes"], ["1500000", "rep"], ["50000", "zeros"], ["300007", "abra"]])
def test_distributed_sa_single_rank(args):
    out = _run(1, args + ["--no-warmup"])
    assert out["parity_vs_single_gpu"] is True, out


def test_distributed_sa_two_ranks():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    for args in (["22", "dna"], ["1500000", "rep"], ["50000", "zeros"]):
        out = _run(2, args + ["--no-warmup"])
        assert out["parity_vs_single_gpu"] is True and out["world"] == 2, out
    out = _run(2, ["24", "dna", "--no-warmup", "--verify-dist"])
    assert out["distributed_check"]["ok"] is True, out
