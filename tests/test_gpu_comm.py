"""
-m gpu tests of the data-parallel exchange (tensorforth_b200/csrc/comm.cu; include/t4k.h "data-parallel extras").

One GPU is enough to exercise the whole protocol: `world` communicators are created in THIS process on cuda:0, wired
with t4k_comm_connect_local, and each rank's kernel is launched on its own stream — the kernels spin on one another's
flags exactly as they do across GPUs (there the pointers are cudaIpc mappings and the stores travel over NVLink).
Checked: SUM all-reduce bit-exact against a rank-ordered sum, mixed call lengths (epoch / parity bookkeeping),
the fused exchange+optimizer against t4k_optim_multi on the summed gradient (bit-exact, every optimizer kind), scalars
riding along, CUDA-graph replay.  The cross-process cudaIpc rendezvous is covered by bench.py --gpus N (NCCL box) and,
for the host logic, by tests/test_dp_gloo.py.
"""
import ctypes as C

import numpy as np
import pytest
import torch

from tensorforth_b200 import lib as t4
from gpu_util import lib, ok, ptr

pytestmark = pytest.mark.gpu


class Ring:
    """`world` ranks of one process on cuda:0"""

    def __init__(self, world, cap):
        L = lib()
        self.L, self.world = L, world
        self.h = (C.c_void_p * world)()
        for r in range(world):
            h = C.c_void_p()
            ok(L.t4k_comm_create(r, world, cap, C.byref(h), None), "comm_create")
            self.h[r] = h
        for r in range(world):
            ok(L.t4k_comm_connect_local(self.h[r], self.h), "connect_local")
        self.streams = [torch.cuda.Stream() for _ in range(world)]

    def st(self, r):
        return C.c_void_p(self.streams[r].cuda_stream)

    def sync(self):
        for s in self.streams:
            s.synchronize()
        for r in range(self.world):
            assert self.L.t4k_comm_status(self.h[r]) == 0, "rank %d: exchange wait timed out" % r

    def close(self):
        torch.cuda.synchronize()
        for r in range(self.world):
            self.L.t4k_comm_destroy(self.h[r])


def ranked_sum(parts):
    s = parts[0].clone()
    for p in parts[1:]:
        s = s + p                       # rank order, FP32: what the kernel does
    return s


def warm(L):
    """load every kernel variant with a world=1 communicator first (a lazy module load while a peer kernel spins would
    serialise behind it)"""
    r1 = Ring(1, 4096)
    b = torch.ones(4096, device="cuda")
    ok(L.t4k_allreduce_sum(r1.h[0], ptr(b), 4096, r1.st(0)))
    ok(L.t4k_allreduce_sum(r1.h[0], ptr(b), 4095, r1.st(0)))
    seg = seg_table([(0, 4096, 1)])
    for kind in (0, 1, 2):
        g, dg, m, v = (torch.ones(4096, device="cuda") for _ in range(4))
        ok(L.t4k_optim_multi_dp(r1.h[0], kind, ptr(g), ptr(dg), ptr(m), ptr(v), ptr(seg), 1, 4096, 1e-3, 0.9, 0.999, 0.0, None, 0, 0, r1.st(0)))
    dg = torch.ones(4096, device="cuda")
    assert L.t4k_dp_push(r1.h[0], ptr(dg), 8, 4096, r1.st(0)) >= 0
    ok(L.t4k_optim_multi_dp(r1.h[0], 1, ptr(g), ptr(dg), ptr(m), ptr(v), ptr(seg), 1, 4096, 1e-3, 0.9, 0.999, 0.0, None, 0, 0, r1.st(0)))
    r1.sync()
    assert torch.equal(b, torch.ones_like(b))           # world=1: the sum over ranks is the identity
    r1.close()


def warm_bn(L):
    """the data-parallel batch-norm kernels, loaded with a world=1 communicator (same reason as warm(): the first launch of a kernel loads its
    module, and a load issued while a peer's kernel spins waits for it — in ONE process that peer's partner is launched by the same thread)"""
    r1 = Ring(1, 64)
    x = torch.ones(2, 4, 8, device="cuda")
    o, xh, dx, s3 = torch.zeros_like(x), torch.zeros_like(x), torch.zeros_like(x), torch.zeros(3 * 8 + 4, device="cuda")
    g, b, dg, db = (torch.ones(8, device="cuda") for _ in range(4))
    ok(L.t4k_batchnorm_fwd_dp(r1.h[0], ptr(x), ptr(o), ptr(xh), ptr(g), ptr(b), ptr(s3), 2, 2, 4, 8, r1.st(0)))
    ok(L.t4k_batchnorm_bwd_dp(r1.h[0], ptr(x), ptr(xh), ptr(dx), ptr(g), ptr(dg), ptr(db), ptr(s3), 2, 2, 4, 8, 1, r1.st(0)))
    x7 = torch.ones(2, 4, 7, device="cuda")                  # the unaligned variants
    o7, xh7, dx7 = torch.zeros_like(x7), torch.zeros_like(x7), torch.zeros_like(x7)
    ok(L.t4k_batchnorm_fwd_dp(r1.h[0], ptr(x7), ptr(o7), ptr(xh7), ptr(g), ptr(b), ptr(s3), 2, 2, 4, 7, r1.st(0)))
    ok(L.t4k_batchnorm_bwd_dp(r1.h[0], ptr(x7), ptr(xh7), ptr(dx7), ptr(g), ptr(dg), ptr(db), ptr(s3), 2, 2, 4, 7, 1, r1.st(0)))
    r1.sync()
    r1.close()


def seg_table(segs):
    a = np.zeros((len(segs), 3), dtype=np.int64)        # {int64 off; int64 len; int32 Nw; int32 pad}
    for k, (off, ln, nw) in enumerate(segs):
        a[k, 0], a[k, 1] = off, ln
        a[k, 2] = nw                                    # little endian: Nw in the low half, pad = 0
    return torch.from_numpy(a).cuda()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_allreduce_sum_mixed_lengths(world):
    L = lib()
    warm(L)
    cap = 480000 // world        # all `world` kernels must be co-resident on this ONE GPU (they spin on one another)
    ring = Ring(world, cap)
    g = torch.Generator(device="cuda").manual_seed(world)
    for n in (cap, 64, 7, cap - 4, cap, 1, 4099, cap // 2 + 3, cap):    # full, tiny, unaligned, repeated
        parts = [torch.randn(n, device="cuda", generator=g) for _ in range(world)]
        want = ranked_sum(parts)
        bufs = [p.clone() for p in parts]
        torch.cuda.synchronize()
        for r in range(world):
            ok(L.t4k_allreduce_sum(ring.h[r], ptr(bufs[r]), n, ring.st(r)), "allreduce")
        ring.sync()
        for r in range(world):
            assert torch.equal(bufs[r], want), "n=%d rank %d" % (n, r)
    ring.close()


@pytest.mark.parametrize("split", [False, True, "range"])
@pytest.mark.parametrize("world,big", [(4, 48000), (2, 196000)])
@pytest.mark.parametrize("kind", [0, 1, 2])
def test_fused_exchange_optimizer_equals_optimizer_on_summed_gradient(kind, world, big, split):
    """split: the exchange in two launches — t4k_dp_push of everything past the first parameter layer's segments (what
    Model::step_graph forks onto a side stream while the first layer's backward still runs), then the fused kernel"""
    L = lib()
    warm(L)
    segs = [(0, 92, 1), (92, 12, 1), (104, big, 1), (big + 104, 100, 1), (big + 204, 1000, 3), (big + 1204, 12, 1)]
    total = big + 1216
    seg = seg_table(segs)
    ring = Ring(world, total)
    gen = torch.Generator(device="cuda").manual_seed(7 + kind)
    G0 = torch.randn(total, device="cuda", generator=gen) * 0.1
    M0 = torch.randn(total, device="cuda", generator=gen) * 0.01
    V0 = torch.rand(total, device="cuda", generator=gen) * 1e-3
    lr, b1, b2, wd = 1e-2, 0.9, 0.999, 1e-3
    st0 = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    G = [G0.clone() for _ in range(world)]; M = [M0.clone() for _ in range(world)]; V = [V0.clone() for _ in range(world)]
    Gr, Mr, Vr = G0.clone(), M0.clone(), V0.clone()
    for step in range(3):
        DG = [torch.randn(total, device="cuda", generator=gen) for _ in range(world)]
        scal = [torch.tensor([1.0 + r, 0.5 * r, step], device="cuda") for r in range(world)]
        DGr = ranked_sum(DG)
        ok(L.t4k_optim_multi(kind, ptr(Gr), ptr(DGr), ptr(Mr), ptr(Vr), ptr(seg), len(segs), total, lr, b1, b2, wd, st0), "optim_multi")
        torch.cuda.synchronize()
        pushed = [0] * world
        if split == "range":
            # the exchange + optimizer in two launches on disjoint chunk ranges (what Model::step_graph does: everything past the first layer's
            # chunk early on a side stream, the first chunk — and the scalars — at the end)
            chf = L.t4k_comm_chunk_floats(ring.h[0])
            cut = ((104 + chf - 1) // chf) * chf
            assert 0 < cut < total
            for r in range(world):
                ok(L.t4k_optim_multi_dp_range(ring.h[r], kind, ptr(G[r]), ptr(DG[r]), ptr(M[r]), ptr(V[r]), ptr(seg), len(segs), cut, total, total,
                                              lr, b1, b2, wd, None, 0, 0, ring.st(r)), "range (rest)")
            for r in range(world):
                ok(L.t4k_optim_multi_dp_range(ring.h[r], kind, ptr(G[r]), ptr(DG[r]), ptr(M[r]), ptr(V[r]), ptr(seg), len(segs), 0, cut, total,
                                              lr, b1, b2, wd, ptr(scal[r]), 3, 0, ring.st(r)), "range (first)")
        elif split:
            for r in range(world):
                pushed[r] = L.t4k_dp_push(ring.h[r], ptr(DG[r]), 104, total, ring.st(r))
                assert 104 <= pushed[r] < total, pushed[r]
        for r in range(world):
            if split == "range":
                break
            ok(L.t4k_optim_multi_dp(ring.h[r], kind, ptr(G[r]), ptr(DG[r]), ptr(M[r]), ptr(V[r]), ptr(seg), len(segs), total,
                                    lr, b1, b2, wd, ptr(scal[r]), 3, pushed[r], ring.st(r)), "optim_multi_dp")
        ring.sync()
        for r in range(world):
            assert torch.equal(G[r], Gr), "G step %d rank %d" % (step, r)
            assert torch.equal(M[r], Mr) and (kind == 0 or torch.equal(V[r], Vr))
            assert float(DG[r].abs().max()) == 0.0                  # consumed, as Model::gradient zeroes dw/db
            assert scal[r].tolist() == [sum(1.0 + k for k in range(world)), sum(0.5 * k for k in range(world)), float(step * world)]
    ring.close()


@pytest.mark.parametrize("world,big", [(2, 196000), (4, 48000)])
def test_dma_push_then_fused_exchange(world, big):
    """the copy-engine push (t4k_dp_push_dma: peer-to-peer copies into the slots of the stated parity + a one-block signal kernel) followed by the
    fused exchange + optimizer — in one launch, or split at the pushed offset into the rest of the arena and the first chunk (what the captured
    step does) — equals the optimizer on the rank-summed gradient, step after step (the parity alternates)"""
    L = lib()
    warm(L)
    segs = [(0, 92, 1), (92, 12, 1), (104, big, 1), (big + 104, 100, 1), (big + 204, 1000, 3), (big + 1204, 12, 1)]
    total = big + 1216
    seg = seg_table(segs)
    ring = Ring(world, total)
    gen = torch.Generator(device="cuda").manual_seed(11)
    G0, M0, V0 = torch.randn(total, device="cuda", generator=gen) * 0.1, torch.randn(total, device="cuda", generator=gen) * 0.01, torch.rand(total, device="cuda", generator=gen) * 1e-3
    lr, b1, b2, wd = 1e-2, 0.9, 0.999, 0.0
    st0 = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    G = [G0.clone() for _ in range(world)]; M = [M0.clone() for _ in range(world)]; V = [V0.clone() for _ in range(world)]
    Gr, Mr, Vr = G0.clone(), M0.clone(), V0.clone()
    for step in range(4):
        DG = [torch.randn(total, device="cuda", generator=gen) for _ in range(world)]
        DGr = ranked_sum(DG)
        ok(L.t4k_optim_multi(1, ptr(Gr), ptr(DGr), ptr(Mr), ptr(Vr), ptr(seg), len(segs), total, lr, b1, b2, wd, st0), "optim_multi")
        torch.cuda.synchronize()
        pushed = [0] * world
        for r in range(world):
            pushed[r] = L.t4k_dp_push_dma(ring.h[r], ptr(DG[r]), 104, total, step, ring.st(r))
            assert 104 <= pushed[r] < total, pushed[r]
        for r in range(world):
            if step & 1:                                            # split at the pushed offset: rest of the arena, then the first chunk
                ok(L.t4k_optim_multi_dp_range(ring.h[r], 1, ptr(G[r]), ptr(DG[r]), ptr(M[r]), ptr(V[r]), ptr(seg), len(segs), pushed[r], total, total,
                                              lr, b1, b2, wd, None, 0, pushed[r], ring.st(r)), "rest")
                ok(L.t4k_optim_multi_dp_range(ring.h[r], 1, ptr(G[r]), ptr(DG[r]), ptr(M[r]), ptr(V[r]), ptr(seg), len(segs), 0, pushed[r], total,
                                              lr, b1, b2, wd, None, 0, pushed[r], ring.st(r)), "first chunk")
            else:
                ok(L.t4k_optim_multi_dp(ring.h[r], 1, ptr(G[r]), ptr(DG[r]), ptr(M[r]), ptr(V[r]), ptr(seg), len(segs), total,
                                        lr, b1, b2, wd, None, 0, pushed[r], ring.st(r)), "optim_multi_dp")
        ring.sync()
        for r in range(world):
            assert torch.equal(G[r], Gr), "G step %d rank %d" % (step, r)
            assert torch.equal(M[r], Mr) and torch.equal(V[r], Vr)
            assert float(DG[r].abs().max()) == 0.0
    # a push addressed with the wrong parity is caught on the device (sticky error), not summed
    DGx = [torch.ones(total, device="cuda") for _ in range(world)]
    assert L.t4k_dp_push_dma(ring.h[0], ptr(DGx[0]), 104, total, 5, ring.st(0)) >= 104         # 4 exchanges completed: step 5 is the wrong parity
    ring.streams[0].synchronize()
    assert L.t4k_comm_status(ring.h[0]) == 1
    torch.cuda.synchronize()
    for r in range(world):
        L.t4k_comm_destroy(ring.h[r])


@pytest.mark.parametrize("kind", [0, 1, 2])
@pytest.mark.parametrize("world,big", [(4, 48000), (8, 20000), (3, 30000)])
def test_reduce_scatter_exchange(kind, world, big):
    """world > 2: every chunk past the first goes to ONE owner (t4k_dp_push_owner), the owners sum in rank order and send the sums to every rank,
    every rank runs the optimizer (t4k_optim_multi_dp_rs); the first chunk keeps the one-hop all-to-all (t4k_optim_multi_dp_range).  Result: the
    optimizer on the rank-summed gradient, bit for bit, replicas identical, step after step."""
    L = lib()
    warm(L)
    segs = [(0, 92, 1), (92, 12, 1), (104, big, 1), (big + 104, 100, 1), (big + 204, 1000, 3), (big + 1204, 12, 1)]
    total = big + 1216
    seg = seg_table(segs)
    ring = Ring(world, total)
    # load the reduce-scatter kernels with nobody waiting (see warm())
    r1 = Ring(1, total)
    w_ = [torch.ones(total, device="cuda") for _ in range(4)]
    assert L.t4k_dp_push_owner(r1.h[0], ptr(w_[1]), 104, total, r1.st(0)) >= 104
    ok(L.t4k_optim_multi_dp_rs(r1.h[0], kind, ptr(w_[0]), ptr(w_[1]), ptr(w_[2]), ptr(w_[3]), ptr(seg), len(segs), L.t4k_comm_chunk_floats(r1.h[0]), total, 1e-3, 0.9, 0.999, 0.0, 0, r1.st(0)))
    ok(L.t4k_optim_multi_dp_range(r1.h[0], kind, ptr(w_[0]), ptr(w_[1]), ptr(w_[2]), ptr(w_[3]), ptr(seg), len(segs), 0, L.t4k_comm_chunk_floats(r1.h[0]), total, 1e-3, 0.9, 0.999, 0.0, None, 0, 0, r1.st(0)))
    r1.sync(); r1.close()
    gen = torch.Generator(device="cuda").manual_seed(13 + kind)
    G0, M0, V0 = torch.randn(total, device="cuda", generator=gen) * 0.1, torch.randn(total, device="cuda", generator=gen) * 0.01, torch.rand(total, device="cuda", generator=gen) * 1e-3
    lr, b1, b2, wd = 1e-2, 0.9, 0.999, 1e-3
    st0 = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    G = [G0.clone() for _ in range(world)]; M = [M0.clone() for _ in range(world)]; V = [V0.clone() for _ in range(world)]
    Gr, Mr, Vr = G0.clone(), M0.clone(), V0.clone()
    for step in range(3):
        DG = [torch.randn(total, device="cuda", generator=gen) for _ in range(world)]
        scal = [torch.tensor([1.0 + r], device="cuda") for r in range(world)]
        DGr = ranked_sum(DG)
        ok(L.t4k_optim_multi(kind, ptr(Gr), ptr(DGr), ptr(Mr), ptr(Vr), ptr(seg), len(segs), total, lr, b1, b2, wd, st0), "optim_multi")
        torch.cuda.synchronize()
        cut = [0] * world
        for r in range(world):
            cut[r] = L.t4k_dp_push_owner(ring.h[r], ptr(DG[r]), 104, total, ring.st(r))
            assert 104 <= cut[r] < total
        if step == 1:                                               # one launch per rank (phase 0)
            for r in range(world):
                ok(L.t4k_optim_multi_dp_rs(ring.h[r], kind, ptr(G[r]), ptr(DG[r]), ptr(M[r]), ptr(V[r]), ptr(seg), len(segs), cut[r], total, lr, b1, b2, wd, 0, ring.st(r)), "rs")
        else:                                                       # the owners' half, then everybody's half (what the captured step does)
            for ph in (1, 2):
                for r in range(world):
                    ok(L.t4k_optim_multi_dp_rs(ring.h[r], kind, ptr(G[r]), ptr(DG[r]), ptr(M[r]), ptr(V[r]), ptr(seg), len(segs), cut[r], total, lr, b1, b2, wd, ph, ring.st(r)), "rs phase %d" % ph)
        for r in range(world):
            ok(L.t4k_optim_multi_dp_range(ring.h[r], kind, ptr(G[r]), ptr(DG[r]), ptr(M[r]), ptr(V[r]), ptr(seg), len(segs), 0, cut[r], total,
                                          lr, b1, b2, wd, ptr(scal[r]), 1, cut[r], ring.st(r)), "first chunk")
        ring.sync()
        for r in range(world):
            assert torch.equal(G[r], Gr), "G step %d rank %d" % (step, r)
            assert torch.equal(M[r], Mr) and (kind == 0 or torch.equal(V[r], Vr))
            assert float(DG[r].abs().max()) == 0.0
            assert scal[r].tolist() == [sum(1.0 + k for k in range(world))]
    ring.close()


def test_exchange_replays_inside_cuda_graphs():
    L = lib()
    warm(L)
    world, n = 2, 50000
    ring = Ring(world, n)
    bufs = [torch.zeros(n, device="cuda") for _ in range(world)]
    src = [torch.full((n,), float(r + 1), device="cuda") for r in range(world)]
    graphs = []
    for r in range(world):                                          # each rank: buf = src; all-reduce(buf)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=ring.streams[r]):
            bufs[r].copy_(src[r])
            ok(L.t4k_allreduce_sum(ring.h[r], ptr(bufs[r]), n, ring.st(r)))
        graphs.append(g)
    for it in range(5):
        for r in range(world):
            src[r].fill_(float(r + 1 + it))
        torch.cuda.synchronize()
        for r in range(world):
            with torch.cuda.stream(ring.streams[r]):
                graphs[r].replay()
        ring.sync()
        for r in range(world):
            assert float(bufs[r].min()) == float(bufs[r].max()) == float(sum(k + 1 + it for k in range(world)))
    ring.close()


# ------------------------------------------------------------------ data-parallel semantics beyond the gradient sum
def test_rand_sharded_is_a_slice_of_the_single_device_draw():
    """per-rank Philox offsets = global element index: the shards drawn by `world` ranks are the slices of ONE draw of the whole tensor,
    and every rank's stream advances by the global length (the next draws stay aligned)"""
    L = lib()
    n, world = 4096 * 3 + 8, 4                               # per-shard elements
    for opt in (t4.UNIFORM, t4.NORMAL):
        L.t4k_rand_seed(99)
        full, nxt = torch.zeros(n * world, device="cuda"), torch.zeros(100, device="cuda")
        ok(L.t4k_rand(ptr(full), n * world, opt, 0.0, 1.0, None)); ok(L.t4k_rand(ptr(nxt), 100, opt, 0.0, 1.0, None))
        torch.cuda.synchronize()
        for r in range(world):
            L.t4k_rand_seed(99)
            sh, nx = torch.zeros(n, device="cuda"), torch.zeros(100, device="cuda")
            ok(L.t4k_rand_sharded(ptr(sh), n, r * n, world * n, opt, 0.0, 1.0, None)); ok(L.t4k_rand(ptr(nx), 100, opt, 0.0, 1.0, None))
            torch.cuda.synchronize()
            assert torch.equal(sh, full[r * n:(r + 1) * n]), (opt, r)
            assert torch.equal(nx, nxt)
    assert L.t4k_rand_sharded(ptr(full), n, 4, n, t4.UNIFORM, 0.0, 1.0, None) == t4.EINVAL        # shard past the end of the tensor


@pytest.mark.parametrize("world,N,HW,Cc", [(2, 8, 16, 6), (4, 4, 49, 32), (2, 16, 1, 100)])
def test_batchnorm_dp_equals_single_device_on_the_whole_batch(world, N, HW, Cc):
    """batch norm over a batch sharded across `world` ranks (statistics SUM-all-reduced over peer memory between the two passes) against
    t4k_batchnorm_fwd / _bwd on the concatenated batch: activations, x-hat, dX within FP32 summation noise; the ranks' dgamma / dbeta
    shares add up to the single-device parameter gradients (the gradient exchange performs that sum)"""
    from oracle import oracle as orc
    L = lib()
    warm(L); warm_bn(L)
    ring = Ring(world, 4 * Cc)
    g = torch.Generator(device="cuda").manual_seed(5)
    X = torch.randn(world * N, HW, Cc, device="cuda", generator=g) * 2 + 0.5
    dY = torch.randn(world * N, HW, Cc, device="cuda", generator=g)
    gamma, beta = torch.rand(Cc, device="cuda", generator=g) + 0.5, torch.randn(Cc, device="cuda", generator=g)
    # single device, whole batch
    O1, XH1, dX1, s1 = torch.zeros_like(X), torch.zeros_like(X), torch.zeros_like(X), torch.zeros(3 * Cc + 4, device="cuda")
    dg1, db1 = torch.ones(Cc, device="cuda"), torch.ones(Cc, device="cuda")
    ok(L.t4k_batchnorm_fwd(ptr(X), ptr(O1), ptr(XH1), ptr(gamma), ptr(beta), ptr(s1), world * N, HW, Cc, None))
    ok(L.t4k_batchnorm_bwd(ptr(dY), ptr(XH1), ptr(dX1), ptr(gamma), ptr(dg1), ptr(db1), ptr(s1), world * N, HW, Cc, 1, None))
    torch.cuda.synchronize()
    # sharded
    Xs, dYs = [X[r * N:(r + 1) * N].contiguous() for r in range(world)], [dY[r * N:(r + 1) * N].contiguous() for r in range(world)]
    Os, XHs, dXs = ([torch.zeros_like(Xs[0]) for _ in range(world)] for _ in range(3))
    ss = [torch.zeros(3 * Cc + 4, device="cuda") for _ in range(world)]
    dgs, dbs = [torch.zeros(Cc, device="cuda") for _ in range(world)], [torch.zeros(Cc, device="cuda") for _ in range(world)]
    # ranks of ONE process: each takes its own bank of the library workspaces (separate processes have their own)
    for r in range(world):
        L.t4k_set_workspace_bank(2 * (r % 4))
        ok(L.t4k_batchnorm_fwd_dp(ring.h[r], ptr(Xs[r]), ptr(Os[r]), ptr(XHs[r]), ptr(gamma), ptr(beta), ptr(ss[r]), N, world * N, HW, Cc, ring.st(r)), "bn fwd dp")
    ring.sync()
    for r in range(world):
        L.t4k_set_workspace_bank(2 * (r % 4))
        ok(L.t4k_batchnorm_bwd_dp(ring.h[r], ptr(dYs[r]), ptr(XHs[r]), ptr(dXs[r]), ptr(gamma), ptr(dgs[r]), ptr(dbs[r]), ptr(ss[r]), N, world * N, HW, Cc, 1, ring.st(r)), "bn bwd dp")
    ring.sync()
    L.t4k_set_workspace_bank(0)
    from gpu_util import assert_close
    assert_close(torch.cat(Os).cpu().numpy(), O1.cpu().numpy(), rtol=1e-5, what="O")
    assert_close(torch.cat(XHs).cpu().numpy(), XH1.cpu().numpy(), rtol=1e-5, what="x-hat")
    assert_close(torch.cat(dXs).cpu().numpy(), dX1.cpu().numpy(), rtol=1e-5, what="dX")
    assert_close(sum(dgs).cpu().numpy(), (dg1 - 1).cpu().numpy(), rtol=1e-5, atol=1e-6, what="dgamma shares")
    assert_close(sum(dbs).cpu().numpy(), (db1 - 1).cpu().numpy(), rtol=1e-5, atol=1e-6, what="dbeta shares")
    for r in range(1, world):                                  # every rank derived the same statistics (rank-ordered sums: same bits)
        assert torch.equal(ss[r][:3 * Cc], ss[0][:3 * Cc])
    # and against the oracle's restatement of k_batchnorm_1/2/3 on the whole batch
    o_ref, xh_ref, _, rvar = orc.batchnorm(X.cpu().numpy().reshape(world * N, HW, 1, Cc), gamma.cpu().numpy(), beta.cpu().numpy())
    assert_close(torch.cat(Os).cpu().numpy(), o_ref, rtol=1e-4, what="O vs oracle")
    dx_ref, dw_ref, db_ref = orc.dbatchnorm(dY.cpu().numpy().reshape(world * N, HW, 1, Cc), xh_ref, gamma.cpu().numpy(), rvar, np.zeros(Cc, np.float32), np.zeros(Cc, np.float32))
    assert_close(torch.cat(dXs).cpu().numpy(), dx_ref, rtol=1e-4, what="dX vs oracle")
    assert_close(sum(dgs).cpu().numpy(), dw_ref, rtol=1e-4, atol=1e-6, what="dgamma vs oracle")
    assert_close(sum(dbs).cpu().numpy(), db_ref, rtol=1e-4, atol=1e-6, what="dbeta vs oracle")
    ring.close()
