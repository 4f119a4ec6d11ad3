"""Multi-GPU data parallel parity (needs >= 2 GPUs on the box; skipped otherwise): the 2-rank step on half batches —
gradient exchange fused into the optimizer kernel over NVLink peer memory, and the NCCL arm — must follow the single-GPU
full-batch loss trajectory (north star: <= 1e-4 loss deviation), and the replicas must stay bit-identical."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.parametrize("kind,batch", [("mnist", 512), ("bn", 64)])
def test_two_gpu_step_follows_single_gpu_trajectory(kind, batch):
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs 2 GPUs")
    world = 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", "29611", os.path.join(ROOT, "tests", "dp_parity_worker.py")]
    env = dict(os.environ, DP_LR="1e-4", DP_MODEL=kind, DP_BATCH=str(batch))
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    line = [l for l in p.stdout.splitlines() if l.startswith("DP_PARITY ")]
    assert line, p.stdout[-2000:] + p.stderr[-3000:]
    r = json.loads(line[-1][len("DP_PARITY "):])
    for arm in ("fused", "nccl"):
        assert r[arm]["status"] == 0
        assert r[arm]["replicas_bit_identical"], arm
        assert r[arm]["loss_max_rel_dev"] <= 1e-4, (arm, r[arm], r["single"])
