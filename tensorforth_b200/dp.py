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


def gather_bytes(payload, world=None):
    """every rank contributes `payload` (bytes of one fixed length); returns the list in rank order.  Works on any
    backend: the bytes travel as a uint8 tensor (on the current CUDA device under NCCL, on the host under gloo)."""
    if not active():
        return [bytes(payload)]
    world = world or dist.get_world_size()
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    mine = torch.tensor(list(payload), dtype=torch.uint8, device=dev)
    outs = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(outs, mine)
    return [bytes(o.cpu().tolist()) for o in outs]


class PeerComm:
    """The exchange path of include/t4k.h's data-parallel extras: one exchange block per rank in its own HBM, mapped
    into every peer with cudaIpc, written by NVLink peer stores from inside the kernels (csrc/comm.cu).  The only job of
    the host is the rendezvous: create, gather the 64-byte handles (collective), connect."""

    def __init__(self, cap_floats, rank=None, world=None):
        import ctypes as C
        from . import lib as _k
        self._k, self._L = _k, _k.load()
        self.rank = dist.get_rank() if rank is None and active() else (rank or 0)
        self.world = dist.get_world_size() if world is None and active() else (world or 1)
        h = C.c_void_p()
        hb = (C.c_char * _k.COMM_HANDLE_BYTES)()
        _k.check(self._L.t4k_comm_create(self.rank, self.world, int(cap_floats), C.byref(h), hb), "t4k_comm_create")
        self.handle = h
        allh = b"".join(gather_bytes(bytes(hb.raw), self.world))
        _k.check(self._L.t4k_comm_connect(self.handle, allh), "t4k_comm_connect")
        if active():
            dist.barrier()                    # nobody pushes before every rank has mapped every block

    def allreduce_sum_(self, flat):
        """in-place SUM over the ranks of a float32 CUDA tensor, on the current torch stream"""
        assert flat.is_cuda and flat.dtype == torch.float32 and flat.is_contiguous()
        import ctypes as C
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        self._k.check(self._L.t4k_allreduce_sum(self.handle, C.c_void_p(flat.data_ptr()), flat.numel(), st), "t4k_allreduce_sum")
        return flat

    def status(self):
        return int(self._L.t4k_comm_status(self.handle))

    def close(self):
        if self.handle:
            self._L.t4k_comm_destroy(self.handle)
            self.handle = None


class DataParallel:
    """Wraps a tensorforth_b200.host.Model that holds this rank's shard of the batch.

        dp = DataParallel(model, device)          # after the model is built (collective: broadcasts parameters)
        model.forward(x_shard); model.backprop(y_shard)
        dp.allreduce_grads()                      # SUM over ranks of the whole DG arena, one collective
        model.adam(lr)                            # identical update on every rank
    """

    def __init__(self, model, device, sync_params=True, fused=False, scalars=None):
        """fused=True: attach a PeerComm to the model — the SUM of the gradient arena over the ranks then happens inside
        the optimizer kernel (model.adam()/sgd()/step_graph), over NVLink peer memory, and allreduce_grads() must not be
        called; `scalars` (a float32 CUDA tensor of <= 64 elements, e.g. the loss sum) is summed in the same exchange.
        fused=False: the gradient arena goes through torch.distributed (NCCL / gloo) in allreduce_grads()."""
        self.model, self.device = model, device
        g, dg, total = model.arena()              # builds the flat arenas on first use
        self.params = device_view(g, total, device) if device.type == "cuda" else None
        self.grads = device_view(dg, total, device) if device.type == "cuda" else None
        self.total = total
        self.comm = self.stat_comm = None
        if sync_params and self.params is not None:
            if self.params.is_cuda:
                from . import host as _h
                with torch.cuda.stream(torch.cuda.ExternalStream(_h.stream(), device=device)):
                    broadcast_(self.params, 0)
            else:
                broadcast_(self.params, 0)
        if active():
            # this rank's shard of the global batch: dropout masks drawn at the shard's global element offsets, batch-norm statistics summed
            # over the ranks (over NVLink peer memory, inside Model::forward / backprop) — with either gradient exchange
            bn = model.bn_channels() if hasattr(model, "bn_channels") else 0
            if bn and device.type != "cuda":
                raise NotImplementedError("batch-norm statistics are exchanged over CUDA peer memory: no CPU path")
            if bn:
                self.stat_comm = PeerComm(4 * bn)
            if hasattr(model, "dp_shard"):
                model.dp_shard(dist.get_rank(), dist.get_world_size(), self.stat_comm.handle if self.stat_comm else None)
        if fused and active():
            import ctypes as C
            self.comm = PeerComm(total)
            self._scalars = scalars               # keep alive
            model.dp_attach(self.comm.handle, C.c_void_p(scalars.data_ptr()) if scalars is not None else None,
                            scalars.numel() if scalars is not None else 0)

    def allreduce_grads(self):
        """SUM of the gradient arena over the ranks (torch.distributed), ordered after the library stream's backprop and before its optimizer:
        the collective is issued on the library's own stream whatever torch's current stream is"""
        if self.comm is not None:
            raise RuntimeError("fused data parallel: the optimizer call does the exchange")
        if self.grads is not None and self.grads.is_cuda:
            from . import host as _h
            with torch.cuda.stream(torch.cuda.ExternalStream(_h.stream(), device=self.device)):
                return allreduce_sum_(self.grads)
        return allreduce_sum_(self.grads)

    def hit(self):
        (h,) = reduce_scalars([self.model.hit(True)], self.device)
        return int(round(h))
