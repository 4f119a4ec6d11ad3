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


@pytest.mark.parametrize("split", [False, True])
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
        if split:
            for r in range(world):
                pushed[r] = L.t4k_dp_push(ring.h[r], ptr(DG[r]), 104, total, ring.st(r))
                assert 104 <= pushed[r] < total, pushed[r]
        for r in range(world):
            ok(L.t4k_optim_multi_dp(ring.h[r], kind, ptr(G[r]), ptr(DG[r]), ptr(M[r]), ptr(V[r]), ptr(seg), len(segs), total,
                                    lr, b1, b2, wd, ptr(scal[r]), 3, pushed[r], ring.st(r)), "optim_multi_dp")
        ring.sync()
        for r in range(world):
            assert torch.equal(G[r], Gr), "G step %d rank %d" % (step, r)
            assert torch.equal(M[r], Mr) and (kind == 0 or torch.equal(V[r], Vr))
            assert float(DG[r].abs().max()) == 0.0                  # consumed, as Model::gradient zeroes dw/db
            assert scal[r].tolist() == [sum(1.0 + k for k in range(world)), sum(0.5 * k for k in range(world)), float(step * world)]
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
