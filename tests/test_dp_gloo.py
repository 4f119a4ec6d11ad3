"""
Data-parallel host logic (tensorforth_b200/dp.py) on CPU: world_size 2, gloo.  The compute on each rank is the CPU
oracle (test infrastructure) standing in for the CUDA path; what is under test is the sharding + SUM all-reduce
contract of SURVEY.md §8e: a 2-rank step on half batches must reproduce the 1-rank step on the full batch
(the reference's gradients are batch sums, so no rescaling is involved).
"""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tensorforth_b200 import dp                     # noqa: E402


def test_shard_bounds_cover_and_balance():
    for n in (0, 1, 7, 512, 1000):
        for world in (1, 2, 3, 8):
            spans = [dp.shard_bounds(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)
    with pytest.raises(ValueError):
        dp.shard_bounds(8, 2, 2)


def test_single_process_is_noop():
    t = torch.arange(4, dtype=torch.float32)
    assert torch.equal(dp.allreduce_sum_(t.clone()), t) and torch.equal(dp.broadcast_(t.clone()), t)
    assert dp.reduce_scalars([1.5, 2]) == [1.5, 2.0]
    assert dp.global_loss(0.25, 8, 8) == 0.25


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _build(orc, N, seed):
    m = orc.OracleModel(N, 8, 8, 1, seed=seed)
    m.add(orc.L_CONV, 4, 0.5, [3, 1, 1, 1]).add(orc.L_MAXPOOL, 2).add(orc.L_RELU).add(orc.L_FLATTEN)
    m.add(orc.L_LINEAR, 12, 1.0).add(orc.L_RELU).add(orc.L_LINEAR, 5, 1.0).add(orc.L_SOFTMAX)
    return m


def _flat(m, attr):
    return np.concatenate([getattr(t, attr + nm).ravel() for t, nm, _ in m._params()])


def _unflat(m, attr, flat):
    o = 0
    for t, nm, _ in m._params():
        a = getattr(t, attr + nm)
        a[...] = flat[o:o + a.size].reshape(a.shape); o += a.size


def _worker(rank, world, port, N, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import oracle as orc
        rng = np.random.default_rng(5)
        X = (rng.random((N, 8, 8, 1), dtype=np.float32) * 2 - 1).astype(np.float32)
        Y = orc.onehot(rng.integers(0, 5, N), 5)
        lo, hi = dp.shard_bounds(N, world, rank)
        m = _build(orc, hi - lo, seed=100 + rank)                 # deliberately different initial weights per rank
        p = torch.from_numpy(_flat(m, ""))
        dp.broadcast_(p, 0); _unflat(m, "", p.numpy())            # replicas made identical
        losses = []
        for _ in range(3):
            m.forward(X[lo:hi])
            losses.append(dp.global_loss(m.loss(orc.LOSS_CE, Y[lo:hi]), hi - lo, N))
            m.backprop(Y[lo:hi])
            g = torch.from_numpy(_flat(m, "d"))
            dp.allreduce_sum_(g); _unflat(m, "d", g.numpy())
            m.adam(1e-2)
        hits = dp.reduce_scalars([orc.hit(m.forward(X[lo:hi]).output(), Y[lo:hi])])[0]
        if rank == 0:
            q.put((_flat(m, ""), losses, hits))
    finally:
        dist.destroy_process_group()


def _rendezvous_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        got = dp.gather_bytes(bytes([17 * (rank + 1)] * 64))        # what PeerComm does with the cudaIpc handles
        if rank == 1:
            q.put(got)
    finally:
        dist.destroy_process_group()


def test_handle_rendezvous_gathers_in_rank_order():
    assert dp.gather_bytes(b"abc") == [b"abc"]                      # not distributed: own handle only
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_rendezvous_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert got == [bytes([17] * 64), bytes([34] * 64)]


def test_two_rank_step_equals_full_batch_step():
    from oracle import oracle as orc
    N, world = 12, 2
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, N, q)) for r in range(world)]
    for p in procs:
        p.start()
    got_params, got_losses, got_hits = q.get()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    # single-process reference: full batch, rank 0's initial weights
    rng = np.random.default_rng(5)
    X = (rng.random((N, 8, 8, 1), dtype=np.float32) * 2 - 1).astype(np.float32)
    Y = orc.onehot(rng.integers(0, 5, N), 5)
    m = _build(orc, N, seed=100)
    ref_losses = []
    for _ in range(3):
        m.forward(X); ref_losses.append(m.loss(orc.LOSS_CE, Y)); m.backprop(Y); m.adam(1e-2)
    ref_hits = orc.hit(m.forward(X).output(), Y)
    np.testing.assert_allclose(got_losses, ref_losses, rtol=1e-5)
    np.testing.assert_allclose(got_params, _flat(m, ""), rtol=1e-4, atol=1e-6)
    assert int(got_hits) == int(ref_hits)


def test_c_abi_shard_info_matches_the_host_helper():
    import ctypes as C
    from tensorforth_b200 import lib
    L = lib.load()
    for n in (0, 1, 7, 512, 8192):
        for world in (1, 2, 3, 8):
            for r in range(world):
                lo, hi = C.c_int64(), C.c_int64()
                assert L.t4k_shard_info(n, world, r, C.byref(lo), C.byref(hi)) == 0
                assert (lo.value, hi.value) == dp.shard_bounds(n, world, r)
    assert L.t4k_shard_info(8, 2, 2, C.byref(lo), C.byref(hi)) == lib.EINVAL


class _FakeModel:
    """what DataParallel touches of tensorforth_b200.host.Model, on the CPU: arena(), bn_channels(), dp_shard()"""

    def __init__(self, bn):
        self.bn, self.shard = bn, None

    def arena(self):
        return 0, 0, 16

    def bn_channels(self):
        return self.bn

    def dp_shard(self, rank, world, comm_stat=None):
        self.shard = (rank, world, comm_stat)


def _dp_host_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        plain = _FakeModel(0)
        dp.DataParallel(plain, torch.device("cpu"))                 # every model learns which shard of the global batch it holds (dropout masks, BN statistics)
        refused = False
        try:
            dp.DataParallel(_FakeModel(6), torch.device("cpu"))     # batch-norm statistics travel over CUDA peer memory: refused without a device, never silently per-shard
        except NotImplementedError:
            refused = True
        if rank == 1:
            q.put((plain.shard, refused))
    finally:
        dist.destroy_process_group()


def test_data_parallel_tells_the_model_its_shard_and_refuses_batchnorm_without_a_device():
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_dp_host_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    shard, refused = q.get()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert shard == (1, 2, None) and refused
