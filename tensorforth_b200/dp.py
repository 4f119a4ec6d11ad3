"""
dp.py — data-parallel plumbing for Model::forward/backprop.  The reference is single-GPU; its
gradients are batch SUMS (src/nn/backprop.cu:97-103, src/nn/nmath.tcu:277,335 — never divided by N), so sharding
the batch over G ranks and SUM-all-reducing the flat gradient arena between `backprop` and the optimizer
reproduces the single-GPU step up to FP summation order.  One process per GPU, torch.distributed for the wire
(NCCL over NVLink on the GPU box; the same code runs on `gloo` for the CPU tests of the host logic).

Nothing here computes: the payloads are views of buffers the CUDA path (libt4k/libt4host) owns.
"""
import os

import torch
import torch.distributed as dist


def env_rank():
    """(rank, world, local_rank) from the torchrun environment (single process: 0, 1, 0)"""
    return int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))


def shard_bounds(n, world, rank):
    """samples [lo, hi) of a batch of n owned by `rank`: contiguous, sizes differ by at most one, first ranks larger"""
    if world < 1 or not (0 <= rank < world) or n < 0:
        raise ValueError("shard_bounds(n=%d, world=%d, rank=%d)" % (n, world, rank))
    q, r = divmod(n, world)
    lo = rank * q + min(rank, r)
    return lo, lo + q + (1 if rank < r else 0)


def device_view(ptr, numel, device):
    """zero-copy float32 torch view of `numel` floats at device address `ptr` (a libt4host arena)"""
    class _Cai:
        __cuda_array_interface__ = {"shape": (int(numel),), "typestr": "<f4", "data": (int(ptr), False), "version": 3}
    return torch.as_tensor(_Cai(), device=device)


def active():
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def allreduce_sum_(flat, group=None):
    """in-place SUM all-reduce of a flat gradient tensor (no-op when not distributed)"""
    if active():
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    return flat


def broadcast_(flat, src=0, group=None):
    """make every replica's parameters identical to rank `src`'s"""
    if active():
        dist.broadcast(flat, src, group=group)
    return flat


def reduce_scalars(values, device="cpu", op="sum"):
    """all-reduce a few host scalars (loss sums, hit counts); float64 so integer counts stay exact.  Returns floats."""
    t = torch.tensor([float(v) for v in values], dtype=torch.float64, device=device)
    if active():
        dist.all_reduce(t, op=dist.ReduceOp.SUM if op == "sum" else dist.ReduceOp.MAX)
    return [float(x) for x in t.cpu()]


def global_loss(local_loss, n_local, n_global, device="cpu"):
    """Model::loss divides the batch SUM by the local N (src/mu/tensor.cu:289-325): the global-batch loss is
    Σ_r loss_r·N_r / N."""
    (s,) = reduce_scalars([local_loss * n_local], device)
    return s / n_global


class DataParallel:
    """Wraps a tensorforth_b200.host.Model that holds this rank's shard of the batch.

        dp = DataParallel(model, device)          # after the model is built (collective: broadcasts parameters)
        model.forward(x_shard); model.backprop(y_shard)
        dp.allreduce_grads()                      # SUM over ranks of the whole DG arena, one collective
        model.adam(lr)                            # identical update on every rank
    """

    def __init__(self, model, device, sync_params=True):
        self.model, self.device = model, device
        g, dg, total = model.arena()              # builds the flat arenas on first use
        self.params = device_view(g, total, device)
        self.grads = device_view(dg, total, device)
        self.total = total
        if sync_params:
            broadcast_(self.params, 0)

    def allreduce_grads(self):
        return allreduce_sum_(self.grads)

    def hit(self):
        (h,) = reduce_scalars([self.model.hit(True)], self.device)
        return int(round(h))
