"""Worker of tests/test_gpu_dp_multi.py, launched under torch.distributed.run with one rank per GPU.

Every rank: the MNIST CNN on its shard of ONE global batch (same seeded weights everywhere), K Adam steps, data parallel
 (a) with the gradient exchange fused into the optimizer kernel over NVLink peer memory (csrc/comm.cu), whole step = one CUDA graph
 (b) with the NCCL all-reduce arm.
Rank 0 also trains the same model on the FULL batch on its own GPU (no data parallelism): the reference trajectory.
Prints one JSON line (rank 0): max relative deviation of the global loss from the single-GPU trajectory for both arms, whether
the replicas' parameters stayed bit-identical, and the exchange status."""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tensorforth_b200 import dp, host as th, lib as t4      # noqa: E402

NG, K, LR = int(os.environ.get("DP_BATCH", "512")), int(os.environ.get("DP_STEPS", "12")), float(os.environ.get("DP_LR", "1e-3"))
KIND = os.environ.get("DP_MODEL", "mnist")      # "bn": conv -> batchnorm -> relu -> dropout -> avgpool -> flatten -> linear -> softmax: batch statistics of
                                                # the GLOBAL batch and the global batch's dropout masks (per-rank Philox offsets) are what make it follow rank 0's run


def main():
    rank, world, local = dp.env_rank()
    torch.cuda.set_device(local); th.init(local)
    L = t4.load()
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_stream(torch.cuda.ExternalStream(th.stream(), device=local))
    dev = torch.device("cuda", local)
    rng = np.random.default_rng(11)
    shape = (NG, 28, 28, 1) if KIND == "mnist" else (NG, 8, 8, 3)
    Xg = (rng.random(shape, dtype=np.float32) * 2 - 1).astype(np.float32)
    Yg = np.eye(10, dtype=np.float32)[rng.integers(0, 10, NG)]
    lo, hi = dp.shard_bounds(NG, world, rank)
    assert (hi - lo) * world == NG, "equal shards expected"

    def build(n):
        L.t4k_rand_seed(1234)                                   # identical initial weights on every rank / arm
        if KIND == "mnist":
            return th.mnist_cnn(n)
        return th.Model(n, 8, 8, 3).conv2d(0.5, 6).batchnorm().relu().dropout(0.3).avgpool(2).flatten().linear(10).softmax()

    def run(arm):
        n = NG if arm == "single" else hi - lo
        m = build(n)
        X = th.Tensor.from_numpy(Xg if arm == "single" else Xg[lo:hi])
        Y = th.Tensor.tensor(n, 1, 10, 1, Yg if arm == "single" else Yg[lo:hi])
        loss_dev = torch.zeros(8, device="cuda"); lp = C.c_void_p(loss_dev.data_ptr())
        d = None
        if arm == "fused":
            d = dp.DataParallel(m, dev, fused=True, scalars=loss_dev[:1])
        elif arm == "nccl":
            d = dp.DataParallel(m, dev)
        losses = []
        for i in range(K):
            if arm == "nccl":
                t4.check(m.step_graph(X, Y, t4.LOSS_CE, lp, optimizer=-1, lr=LR), "graph")
                d.allreduce_grads(); m.adam(LR)
                losses.append(dp.global_loss(float(loss_dev[0].cpu()), n, NG, dev))
            else:
                t4.check(m.step_graph(X, Y, t4.LOSS_CE, lp, optimizer=2, lr=LR), "graph")
                v = float(loss_dev[0].cpu())
                losses.append(v / world if arm == "fused" else v)
        g, _, total = m.arena()
        params = dp.device_view(g, total, dev).clone()
        st = d.comm.status() if (d is not None and d.comm is not None) else 0
        return losses, params, st

    out = {}
    ref_losses = None
    if rank == 0:
        ref_losses, ref_params, _ = run("single")
    for arm in ("fused", "nccl"):
        dist.barrier()
        losses, params, st = run(arm)
        gathered = [torch.empty_like(params) for _ in range(world)]
        dist.all_gather(gathered, params)
        same = all(torch.equal(gathered[0], g) for g in gathered[1:])
        if rank == 0:
            dev_loss = max(abs(a - b) / max(abs(b), 1e-6) for a, b in zip(losses, ref_losses))
            dpar = float((params - ref_params).abs().max() / ref_params.abs().max())
            out[arm] = {"loss_max_rel_dev": dev_loss, "param_max_rel_dev": dpar, "replicas_bit_identical": bool(same), "status": st,
                        "losses": [round(x, 6) for x in losses]}
    if rank == 0:
        out["single"] = [round(x, 6) for x in ref_losses]
        out["world"] = world
        print("DP_PARITY " + json.dumps(out), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
