"""Multi-GPU (NCCL) equivalence of the CUDA path: 1 rank vs 2 ranks, rays sharded inside an item (all-gather forward /
reduce-scatter backward, SURVEY.md section 8e collective 2) and batch items sharded (collective 1), equal images and equal
gradients.  Needs >= 2 GPUs on the box (`gpurun --gpus 2`); skipped otherwise.  The gloo twin of the collectives runs on CPU in
tests/test_dist_cpu.py."""
import json
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
def test_one_rank_vs_two_ranks(tmp_path):
    out = str(tmp_path / "res.json")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "_dist_gpu_worker.py"), out]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, (r.stdout[-3000:], r.stderr[-3000:])
    res = json.load(open(out))
    print(res)
    assert res["world"] == 2
    assert res["rays"]["image_equal"] and res["rays"]["loss_equal"], res["rays"]       # ray sharding is bit-exact in the forward
    assert res["rays"]["grad_cosine"] >= 0.99999 and res["rays"]["grad_max_rel"] <= 2e-3, res["rays"]   # atomics reorder fp32 sums
    assert res["items"]["grad_cosine"] >= 0.99999 and res["items"]["grad_max_rel"] <= 2e-3, res["items"]
