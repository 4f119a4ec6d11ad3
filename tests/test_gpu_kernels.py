"""
-m gpu parity tests: every entry point of include/t4k.h, called through the C-ABI on cuda:0,
against the CPU oracle (oracle/) on the same seeded inputs.  Bar (north star): bit-exact for
index / reshape / routing work, <= 1e-4 relative for FP32 math (tolerance written per test).
Edge cases follow the reference's own: ragged (non tile-multiple) sizes, n == 0, N-broadcast,
unaligned slices, the four conv (K,S,P) configs, pool K in {2,3}, ties in max-pool routing.
"""
import ctypes as C
import numpy as np
import pytest
import torch

from oracle import oracle as orc
from tensorforth_b200 import lib as t4
from gpu_util import lib, dev, zeros, ptr, host, ok, assert_close, assert_exact

pytestmark = pytest.mark.gpu
RNG = np.random.default_rng(1234)


def rnd(*shape, lo=-1.0, hi=1.0):
    return (RNG.random(shape, dtype=np.float32) * (hi - lo) + lo).astype(np.float32)


# ------------------------------------------------------------------ elementwise
MAP_OPS = [(t4.ABS, 0), (t4.NEG, 0), (t4.EXP, 0), (t4.LN, 0), (t4.LOG, 0), (t4.TANH, 0), (t4.RELU, 0),
           (t4.SIGM, 0), (t4.SQRT, 0), (t4.RCP, 0), (t4.SAT, 0), (t4.FILL, 3.25), (t4.GFILL, 2.0),
           (t4.SCALE, 1.5), (t4.POW, 2.5), (t4.ADD, 0.75), (t4.SUB, 0.75), (t4.MUL, -3.0), (t4.DIV, 7.0)]


@pytest.mark.parametrize("op,v", MAP_OPS)
@pytest.mark.parametrize("n,off", [(1, 0), (1027, 0), (4096, 1), (300001, 0)])
def test_map(op, v, n, off):
    a = rnd(n + off, lo=-2, hi=2)
    if op in (t4.POW, t4.RCP):
        a = np.abs(a) + 0.1
    d = dev(a)
    ok(lib().t4k_map(op, ptr(d, off), v, n, None))
    ref = orc.map_(op, a[off:], v)
    got = host(d)[off:]
    if op in (t4.ABS, t4.NEG, t4.RELU, t4.SAT, t4.FILL, t4.SCALE, t4.ADD, t4.SUB, t4.MUL, t4.DIV, t4.SQRT, t4.RCP):
        assert_exact(got, ref, "map %d" % op)           # IEEE ops: bit exact
    else:
        assert_close(got, ref, rtol=1e-5, atol=1e-6, what="map %d" % op)   # __expf/__logf/__powf vs libm
    if off:
        assert host(d)[0] == a[0]                        # neighbours untouched


def test_map_empty_and_bad():
    d = zeros(4)
    assert lib().t4k_map(t4.ABS, ptr(d), 0.0, 0, None) == 0
    assert lib().t4k_map(t4.IDEN, ptr(d), 0.0, 4, None) != 0      # not a k_math op (t4math.cu:199)
    assert lib().t4k_map(t4.ABS, None, 0.0, 4, None) != 0


@pytest.mark.parametrize("op", [t4.ADD, t4.SUB, t4.MUL, t4.DIV])
def test_ts_tt_ops(op):
    a, b = rnd(5, 333), rnd(5, 333, lo=0.5, hi=2)
    da, db, do = dev(a), dev(b), zeros(5, 333)
    ok(lib().t4k_ts_op(op, ptr(da), 1.7, ptr(do), a.size, None))
    assert_exact(host(do), orc.ts_op(op, a, 1.7))
    ok(lib().t4k_tt_op(op, ptr(da), ptr(db), ptr(do), 333, 5, 5, None))
    assert_exact(host(do), orc.tt_op(op, a, b))
    # N-broadcast of B (tensor.cu:39-46) and of A
    b1 = b[:1]
    ok(lib().t4k_tt_op(op, ptr(da), ptr(dev(b1)), ptr(do), 333, 5, 1, None))
    assert_exact(host(do), orc.tt_op(op, a, b1.reshape(333)))
    ok(lib().t4k_tt_op(op, ptr(dev(b1)), ptr(da), ptr(do), 333, 1, 5, None))
    assert_exact(host(do), orc.tt_op(op, b1.reshape(333), a))
    # vectorised broadcast path (hwc % 4 == 0)
    a4, b4 = rnd(3, 64), rnd(1, 64, lo=0.5, hi=2)
    o4 = zeros(3, 64)
    ok(lib().t4k_tt_op(op, ptr(dev(a4)), ptr(dev(b4)), ptr(o4), 64, 3, 1, None))
    assert_exact(host(o4), orc.tt_op(op, a4, b4.reshape(64)))
    assert lib().t4k_tt_op(op, ptr(da), ptr(db), ptr(do), 333, 5, 2, None) != 0     # N mismatch (tensor.cu:35-38)


@pytest.mark.parametrize("n,off", [(0, 0), (3, 0), (1 << 20, 0), (77777, 3)])
def test_copy(n, off):
    a = rnd(n + off + 8)
    d, o = dev(a), zeros(n + off + 8)
    ok(lib().t4k_copy(ptr(d, off), ptr(o, off), n, None))
    assert_exact(host(o)[off:off + n], a[off:off + n])
    assert not host(o)[off + n:].any()


@pytest.mark.parametrize("N,H,W,Cc", [(1, 2, 3, 1), (2, 70, 33, 1), (1, 512, 1024, 1), (2, 5, 7, 3)])
def test_transpose_identity(N, H, W, Cc):
    a = rnd(N, H, W, Cc)
    o = zeros(N, W, H, Cc)
    ok(lib().t4k_transpose(ptr(dev(a)), ptr(o), N, H, W, Cc, None))
    ref = np.stack([orc.transpose(a[n], Cc) for n in range(N)])
    assert_exact(host(o), ref)
    e = zeros(N, H, W, Cc)
    ok(lib().t4k_identity(ptr(e), N, H, W, Cc, None))
    assert_exact(host(e), np.stack([orc.identity(H, W, Cc) for _ in range(N)]))


# ------------------------------------------------------------------ reductions / losses
@pytest.mark.parametrize("n,off", [(1, 0), (15, 0), (1000, 1), (1 << 22, 0)])
def test_reductions(n, off):
    a = rnd(n + off)
    d, out = dev(a), zeros(4)
    x = a[off:]
    ok(lib().t4k_sum(ptr(d, off), n, ptr(out), None))
    assert_close(host(out)[0], orc.tsum(x), rtol=1e-5, atol=1e-5 * np.sqrt(n))
    ok(lib().t4k_nvar(ptr(d, off), 0.25, n, ptr(out), None))
    assert_close(host(out)[0], float(orc.lib().orc_nvar(orc._p(orc.f32(x)), 0.25, n)), rtol=1e-5)
    ok(lib().t4k_minmax(ptr(d, off), n, 1, ptr(out), None))
    assert host(out)[0] == orc.tmax(x)
    ok(lib().t4k_minmax(ptr(d, off), n, 0, ptr(out), None))
    assert host(out)[0] == orc.tmin(x)
    ok(lib().t4k_avg_std(ptr(d, off), n, ptr(out), None))
    assert_close(host(out)[0], orc.avg(x), rtol=1e-4, atol=1e-6)
    assert_close(host(out)[1], orc.std(x), rtol=1e-4, atol=1e-9)       # reference's std = sqrt(Σ(x-μ)²)/n
    # repeatability: no float atomics → identical bits run to run
    ok(lib().t4k_sum(ptr(d, off), n, ptr(out), None)); s1 = host(out)[0].copy()
    ok(lib().t4k_sum(ptr(d, off), n, ptr(out), None)); s2 = host(out)[0].copy()
    assert s1.tobytes() == s2.tobytes()


def test_dot():
    K, Cc, N = 1000, 3, 4
    a, b = rnd(N, K, Cc), rnd(N, K, Cc)
    o0 = rnd(N, Cc)
    o = dev(o0)
    ok(lib().t4k_dot(ptr(dev(a)), ptr(dev(b)), ptr(o), 0.5, 2.0, K, Cc, N, N, None))
    ref = np.stack([orc.dot(a[n], b[n], o0[n], 0.5, 2.0, Cc) for n in range(N)])
    assert_close(host(o), ref, rtol=1e-5)
    v1, v2 = rnd(100003), rnd(100003)
    o = dev(np.array([3.0], np.float32))
    ok(lib().t4k_dot(ptr(dev(v1)), ptr(dev(v2)), ptr(o), 1.0, 0.0, v1.size, 1, 1, 1, None))
    assert_close(host(o)[0], np.dot(v1.astype(np.float64), v2.astype(np.float64)), rtol=1e-5, atol=1e-3)


@pytest.mark.parametrize("kind", [t4.LOSS_MSE, t4.LOSS_BCE, t4.LOSS_CE, t4.LOSS_NLL])
@pytest.mark.parametrize("N,E", [(1, 2), (3, 2), (512, 10), (1024, 1)])
def test_loss(kind, N, E):
    out = rnd(N, E, lo=0.01, hi=0.99)
    tgt = orc.onehot(RNG.integers(0, E, N), E) if kind != t4.LOSS_MSE else rnd(N, E)
    do = dev(out)
    l = zeros(1)
    ok(lib().t4k_loss(kind, ptr(do), ptr(dev(tgt)), out.size, N, ptr(l), None))
    assert_close(host(l)[0], orc.loss(kind, out, tgt, N), rtol=1e-5, atol=1e-6)
    assert_exact(host(do), out)                          # non-destructive (works on no copy at all)


def test_nan_inf():
    a = rnd(100000)
    a[[5, 77, 9999]] = [np.nan, np.inf, -np.inf]
    cnt = torch.zeros(1, dtype=torch.int32, device="cuda")
    ok(lib().t4k_nan_inf(ptr(dev(a)), a.size, C.c_void_p(cnt.data_ptr()), None))
    assert int(cnt.cpu()[0]) == 3


# ------------------------------------------------------------------ GEMM
def gemm_ref64(A, B, O, alpha, beta, tA, tB):
    a = A.astype(np.float64).T if tA else A.astype(np.float64)
    b = B.astype(np.float64).T if tB else B.astype(np.float64)
    return alpha * (a @ b) + beta * O.astype(np.float64)


@pytest.mark.parametrize("engine", [t4.GEMM_SIMT, t4.GEMM_TC, t4.GEMM_TCF, t4.GEMM_TL])
@pytest.mark.parametrize("tA,tB", [(0, 0), (0, 1), (1, 0), (1, 1)])
@pytest.mark.parametrize("M,N,K", [(2, 2, 3), (64, 64, 64), (128, 128, 32), (200, 100, 70), (67, 63, 45), (130, 260, 513), (512, 100, 1960)])
def test_gemm_engines(engine, tA, tB, M, N, K):
    if engine == t4.GEMM_TL and (((M if tA else K) % 4) or ((K if tB else N) % 4)):
        pytest.skip("layer GEMM: TMA needs 16-byte row pitches (AUTO routes these shapes to the other engines)")
    A = rnd(K, M) if tA else rnd(M, K)
    B = rnd(N, K) if tB else rnd(K, N)
    O0 = rnd(M, N)
    o = dev(O0)
    ok(lib().t4k_gemm_ex(engine, ptr(dev(A)), ptr(dev(B)), ptr(o), 0.5, 2.0, tA, tB, M, N, K, 1, 1, 0, 0, 0, None), "gemm")
    got = host(o)
    ref = orc.gemm(A, B, O0, 0.5, 2.0, bool(tA), bool(tB), M, N, K, 1)      # oracle: FP32 FMA, k ascending (t4math.cu:554-564)
    assert_close(got, ref, rtol=1e-4, what="gemm vs oracle")                 # the north-star bar
    assert_close(got, gemm_ref64(A, B, O0, 0.5, 2.0, tA, tB), rtol=2e-5, what="gemm vs f64")   # and FP32-grade vs exact


def test_gemm_beta0_ignores_garbage():
    M = N = K = 96
    A, B = rnd(M, K), rnd(K, N)
    for eng in (t4.GEMM_SIMT, t4.GEMM_TC, t4.GEMM_TCF, t4.GEMM_TL):
        o = dev(np.full((M, N), np.nan, np.float32))
        ok(lib().t4k_gemm_ex(eng, ptr(dev(A)), ptr(dev(B)), ptr(o), 1.0, 0.0, 0, 0, M, N, K, 1, 1, 0, 0, 0, None))
        assert_close(host(o), gemm_ref64(A, B, np.zeros((M, N)), 1, 0, 0, 0), rtol=2e-5)


def test_gemm_channels_and_batch_broadcast():
    # rank-4 `@`: per-channel, per-sample GEMM with channel stride C; N-broadcast of B (tensor.cu:162-180)
    Nb, M, N, K, Cc = 3, 20, 17, 33, 2
    A, B = rnd(Nb, M, K, Cc), rnd(1, K, N, Cc)
    o = zeros(Nb, M, N, Cc)
    ok(lib().t4k_gemm(ptr(dev(A)), ptr(dev(B)), ptr(o), 1.0, 0.0, 0, 0, M, N, K, Cc, Nb, M * K * Cc, 0, M * N * Cc, None))
    ref = np.stack([orc.gemm(A[n], B[0], None, 1.0, 0.0, False, False, M, N, K, Cc) for n in range(Nb)])
    assert_close(host(o), ref, rtol=1e-5)


def test_gemm_tc_large_vs_f64():
    # 1024^3 on the tcgen05 engine vs exact (float64) product; zero-mean data is the hard case for TF32
    M = N = K = 1024
    A, B = rnd(M, K), rnd(K, N)
    o = zeros(M, N)
    ok(lib().t4k_gemm_ex(t4.GEMM_TC, ptr(dev(A)), ptr(dev(B)), ptr(o), 1.0, 0.0, 0, 0, M, N, K, 1, 1, 0, 0, 0, None))
    ref = A.astype(np.float64) @ B.astype(np.float64)
    got = host(o)
    assert_close(got, ref, rtol=1e-5, what="3xTF32 1024^3")
    rel = np.abs(got - ref).max() / np.abs(ref).max()
    assert rel < 2e-6, rel                                # FP32-grade: plain TF32 would be ~5e-4


@pytest.mark.parametrize("tA,tB,M,N,K", [(0, 1, 1024, 512, 784), (1, 0, 512, 784, 1024), (0, 0, 1024, 784, 512),     # GAN D layer 1: fwd / dW / dX
                                         (0, 1, 512, 100, 1960), (1, 0, 100, 1960, 512), (0, 0, 512, 1960, 100),     # MNIST linear 1960->100
                                         (0, 0, 300, 132, 2052), (1, 1, 129, 257, 36)])                              # ragged tiles, K tails
@pytest.mark.parametrize("engine", [t4.GEMM_TCF, t4.GEMM_AUTO])
def test_gemm_tcf_layer_shapes_vs_f64(engine, tA, tB, M, N, K):
    """the single-launch tensor-core engines (in-kernel 3xTF32 split, split-K: tcgen05 `tcf`, warp-level `mma` with its last-CTA
    finish, and whatever AUTO picks) on the linear-layer shapes: FP32-grade vs exact"""
    A = rnd(K, M) if tA else rnd(M, K)
    B = rnd(N, K) if tB else rnd(K, N)
    O0 = rnd(M, N)
    for alpha, beta in ((1.0, 0.0), (1.0, 1.0)):           # beta = 1: the dW accumulation of Model::_blinear
        o = dev(O0)
        ok(lib().t4k_gemm_ex(engine, ptr(dev(A)), ptr(dev(B)), ptr(o), alpha, beta, tA, tB, M, N, K, 1, 1, 0, 0, 0, None), "gemm layer engine")
        ref = gemm_ref64(A, B, O0, alpha, beta, tA, tB)
        got = host(o)
        assert_close(got, ref, rtol=1e-5, what="tcf vs f64")
        assert np.abs(got - ref).max() / np.abs(ref).max() < 3e-6


TL_SHAPES = [(0, 1, 1024, 512, 784), (1, 0, 512, 784, 1024), (0, 0, 1024, 784, 512),       # GAN D layer 1: fwd / dW / dX
             (0, 1, 512, 100, 1960), (1, 0, 100, 1960, 512), (0, 0, 512, 1960, 100),       # MNIST linear 1960->100
             (0, 1, 1024, 256, 128), (1, 0, 256, 512, 1024), (0, 0, 1024, 256, 512),       # GAN G layers
             (0, 0, 300, 132, 2052), (1, 1, 132, 260, 36), (1, 0, 36, 48, 4100), (0, 1, 64, 16, 4096),   # ragged tiles, K tails, thin outputs
             (1, 1, 256, 128, 3000), (0, 0, 128, 128, 8192)]                                # long K per rank: several accumulator chains


@pytest.mark.parametrize("tA,tB,M,N,K", TL_SHAPES)
def test_gemm_tl_layer_shapes_vs_f64(tA, tB, M, N, K):
    """the layer GEMM (gemm_tl.cu: TMA-fed raw tiles, MN-major operands native, lo plane derived in shared memory, cluster split-K over
    distributed shared memory): FP32-grade vs exact for every transposition, alpha/beta, ragged tiles and K tails"""
    A = rnd(K, M) if tA else rnd(M, K)
    B = rnd(N, K) if tB else rnd(K, N)
    O0 = rnd(M, N)
    for alpha, beta in ((1.0, 0.0), (1.0, 1.0), (0.5, 2.0)):
        o = dev(O0)
        ok(lib().t4k_gemm_ex(t4.GEMM_TL, ptr(dev(A)), ptr(dev(B)), ptr(o), alpha, beta, tA, tB, M, N, K, 1, 1, 0, 0, 0, None), "gemm tl")
        ref = gemm_ref64(A, B, O0, alpha, beta, tA, tB)
        got = host(o)
        assert_close(got, ref, rtol=1e-5, what="tl vs f64")
        assert np.abs(got - ref).max() / np.abs(ref).max() < 3e-6
    assert_close(got, orc.gemm(A, B, O0, 0.5, 2.0, bool(tA), bool(tB), M, N, K, 1), rtol=1e-4, what="tl vs oracle")


@pytest.mark.parametrize("tA,tB,M,N,K", [(0, 1, 512, 100, 1960), (1, 0, 100, 1960, 512), (0, 0, 512, 1960, 100), (1, 1, 132, 260, 36)])
def test_gemm_tl_hi_plane_and_cluster_sizes(tA, tB, M, N, K):
    """(i) the raw FP32 plane as the hi operand (the tensor core ignores the 13 low mantissa bits) must give the same BITS as storing the
    masked hi explicitly; (ii) every cluster size (split-K factor) gives an FP32-grade result, deterministic from call to call"""
    A = rnd(K, M) if tA else rnd(M, K)
    B = rnd(N, K) if tB else rnd(K, N)
    dA, dB_ = dev(A), dev(B)
    ref = gemm_ref64(A, B, np.zeros((M, N)), 1.0, 0.0, tA, tB)
    L = lib()

    def run():
        o = zeros(M, N)
        ok(L.t4k_gemm_ex(t4.GEMM_TL, ptr(dA), ptr(dB_), ptr(o), 1.0, 0.0, tA, tB, M, N, K, 1, 1, 0, 0, 0, None), "gemm tl")
        return host(o).copy()
    base = run()
    assert_exact(run(), base, "deterministic")
    was = L.t4k_set_gemm_tl(1, 1)
    try:
        assert_exact(run(), base, "masked hi == raw plane as hi")
    finally:
        L.t4k_set_gemm_tl(1, was)
    for smax in (1, 2, 4, 8):
        was = L.t4k_set_gemm_tl(2, smax)
        try:
            got = run()
        finally:
            L.t4k_set_gemm_tl(2, was)
        assert_close(got, ref, rtol=1e-5, what="cluster size <= %d" % smax)


@pytest.mark.parametrize("N,E2,EH,E1,act", [(512, 10, 100, 1960, True), (512, 10, 100, 1960, False), (200, 16, 128, 516, True), (64, 3, 20, 64, True)])
def test_linear_dx_from_head(N, E2, EH, E1, act):
    """dX of the hidden linear layer computed straight from the head's forward tensors (p - y, the small linear's dX and the activation
    backward evaluated in the GEMM's operand producer) == t4k_mlp_head_bwd followed by the dX GEMM on its stored output: the same bits"""
    P, T, X2, W2 = np.abs(rnd(N, E2)), orc.onehot(np.arange(N) % E2, E2), rnd(N, EH), rnd(E2, EH)
    F1 = (rnd(N, EH) > 0).astype(np.float32)
    W1 = rnd(EH, E1) * 0.05
    dP, dT, dW2_, dF1, dW1 = dev(P), dev(T), dev(W2), dev(F1), dev(W1)
    # reference path: head backward (writes dY1 into y1 / dX2 into x2), then the dX GEMM
    p, yl, x2, y1 = dev(P), zeros(N, E2), dev(X2), zeros(N, EH)
    dw, db, db1 = zeros(E2, EH), zeros(E2), zeros(EH)
    ok(lib().t4k_mlp_head_bwd(ptr(p), ptr(dT), ptr(yl), ptr(x2), ptr(dF1) if act else None, ptr(y1) if act else None, ptr(dW2_),
                              ptr(dw), ptr(db), ptr(db1), N, E2, EH, 1, None))
    dy1 = y1 if act else x2
    ref = zeros(N, E1)
    ok(lib().t4k_gemm(ptr(dy1), ptr(dW1), ptr(ref), 1.0, 0.0, 0, 0, N, E1, EH, 1, 1, 0, 0, 0, None))
    got = zeros(N, E1)
    rc = lib().t4k_linear_dx_from_head(ptr(dP), ptr(dT), ptr(dW2_), ptr(dF1) if act else None, ptr(dW1), ptr(got), N, E2, EH, E1, None)
    if (N, E2, EH, E1) == (64, 3, 20, 64):
        assert rc in (0, t4.ENOSUP)                        # below the layer GEMM's size class: the caller keeps the two-kernel path
        if rc:
            return
    else:
        ok(rc, "dx_from_head")
    assert_exact(host(got), host(ref), "dX from the head's forward tensors")
    d = orc.tt_op(orc.SUB, P, T)
    g = orc.gemm(d, W2) * (F1 if act else 1.0)
    assert_close(host(got), orc.gemm(np.ascontiguousarray(g, np.float32), W1), rtol=1e-4, what="dX vs oracle")


@pytest.mark.parametrize("N,E2,EH,E1,act", [(512, 10, 100, 1960, True), (512, 10, 100, 1960, False), (200, 16, 128, 516, True), (1024, 10, 64, 784, True)])
def test_linear_bwd_from_head(N, E2, EH, E1, act):
    """dX AND dW of the hidden linear layer in one launch, their dY operand generated from the head's forward tensors (K-major for dX, M-major
    for dW) == t4k_mlp_head_bwd followed by the two GEMMs on its stored output: dX the same bits, dW FP32-grade (its split may differ)"""
    P, T, W2 = np.abs(rnd(N, E2)), orc.onehot(np.arange(N) % E2, E2), rnd(E2, EH)
    X2, X = rnd(N, EH), rnd(N, E1)
    F1 = (rnd(N, EH) > 0).astype(np.float32)
    W1, dW0 = rnd(EH, E1) * 0.05, rnd(EH, E1)
    dP, dT, dW2_, dF1, dW1, dX_ = dev(P), dev(T), dev(W2), dev(F1), dev(W1), dev(X)
    p, yl, x2, y1 = dev(P), zeros(N, E2), dev(X2), zeros(N, EH)
    dw, db, db1 = zeros(E2, EH), zeros(E2), zeros(EH)
    ok(lib().t4k_mlp_head_bwd(ptr(p), ptr(dT), ptr(yl), ptr(x2), ptr(dF1) if act else None, ptr(y1) if act else None, ptr(dW2_),
                              ptr(dw), ptr(db), ptr(db1), N, E2, EH, 1, None))
    dy1 = y1 if act else x2
    ref_dx, ref_dw = zeros(N, E1), dev(dW0)
    ok(lib().t4k_gemm(ptr(dy1), ptr(dW1), ptr(ref_dx), 1.0, 0.0, 0, 0, N, E1, EH, 1, 1, 0, 0, 0, None))
    ok(lib().t4k_gemm(ptr(dy1), ptr(dX_), ptr(ref_dw), 1.0, 1.0, 1, 0, EH, E1, N, 1, 1, 0, 0, 0, None))
    got_dx, got_dw = zeros(N, E1), dev(dW0)
    ok(lib().t4k_linear_bwd_from_head(ptr(dP), ptr(dT), ptr(dW2_), ptr(dF1) if act else None, ptr(dX_), ptr(dW1), ptr(got_dx), ptr(got_dw),
                                      N, E2, EH, E1, None), "bwd_from_head")
    assert_close(host(got_dx), host(ref_dx), rtol=1e-5, what="dX (one launch)")
    assert_close(host(got_dw), host(ref_dw), rtol=1e-5, what="dW (one launch)")
    g = (orc.gemm(orc.tt_op(orc.SUB, P, T), W2) * (F1 if act else 1.0)).astype(np.float32)
    assert_close(host(got_dx), orc.gemm(np.ascontiguousarray(g), W1), rtol=1e-4, what="dX vs oracle")
    assert_close(host(got_dw), orc.gemm(np.ascontiguousarray(g), X, O=dW0, alpha=1.0, beta=1.0, tA=True), rtol=1e-4, what="dW vs oracle")
    a2 = host(got_dw).copy()
    got_dw2, got_dx2 = dev(dW0), zeros(N, E1)
    ok(lib().t4k_linear_bwd_from_head(ptr(dP), ptr(dT), ptr(dW2_), ptr(dF1) if act else None, ptr(dX_), ptr(dW1), ptr(got_dx2), ptr(got_dw2),
                                      N, E2, EH, E1, None), "bwd_from_head")
    assert_exact(host(got_dw2), a2, "deterministic"); assert_exact(host(got_dx2), host(got_dx), "deterministic")


@pytest.mark.parametrize("layer", [t4.L_RELU, t4.L_LEAKYRL, t4.L_TANH])
@pytest.mark.parametrize("N,E1,EH,E0", [(512, 1960, 100, 10), (200, 516, 128, 16), (64, 784, 64, 3)])
def test_linear_act_head_train_equals_fwd_then_head_bwd(layer, N, E1, EH, E0):
    """the TRAIN TAIL (forward tail + the head's backward on the same rows, one launch) leaves the layer tensors exactly as
    t4k_linear_act_head_fwd followed by t4k_mlp_head_bwd leaves them (same bits), and its per-CTA partials add up to that kernel's
    parameter gradients (other summation order: FP32 rounding noise)"""
    X, W1, B1 = rnd(N, E1), rnd(EH, E1) * 0.05, rnd(EH)
    W2, B2 = rnd(E0, EH) * 0.3, rnd(E0)
    T = orc.onehot(np.arange(N) % E0, E0)
    dW20, dB20, dB10 = rnd(E0, EH), rnd(E0), rnd(EH)
    Xd, W1d, B1d, W2d, B2d, Td = dev(X), dev(W1), dev(B1), dev(W2), dev(B2), dev(T)
    L = lib()
    # reference: two launches
    y1a, a1a, f1a, y2a, pa, pda = zeros(N, EH), zeros(N, EH), zeros(N, EH), zeros(N, E0), zeros(N, E0), zeros(N, E0)
    ok(L.t4k_linear_act_head_fwd(layer, ptr(Xd), ptr(W1d), ptr(B1d), ptr(y1a), ptr(a1a), ptr(f1a), 0.1, ptr(W2d), ptr(B2d), ptr(y2a), ptr(pa), ptr(pda),
                                 N, EH, E1, E0, None))
    dwa, dba, db1a = dev(dW20), dev(dB20), dev(dB10)
    ok(L.t4k_mlp_head_bwd(ptr(pa), ptr(Td), ptr(y2a), ptr(a1a), ptr(f1a), ptr(y1a), ptr(W2d), ptr(dwa), ptr(dba), ptr(db1a), N, E0, EH, 1, None))
    # train tail
    nf = L.t4k_head_train_scratch_floats(layer, N, EH, E1, E0)
    if nf == 0:
        pytest.skip("shape outside the train tail's envelope (the caller keeps the two launches)")
    # a forward tail with OTHER head weights in between: the train tail must stage W2 itself (shared memory outlives a kernel: a launch
    # that relied on what the reference launch above left there would still pass without this)
    W2x, junk = dev(rnd(E0, EH) * 3.0), [zeros(N, EH) for _ in range(3)] + [zeros(N, E0) for _ in range(3)]
    ok(L.t4k_linear_act_head_fwd(layer, ptr(Xd), ptr(W1d), ptr(B1d), ptr(junk[0]), ptr(junk[1]), ptr(junk[2]), 0.1, ptr(W2x), ptr(B2d), ptr(junk[3]), ptr(junk[4]),
                                 ptr(junk[5]), N, EH, E1, E0, None))
    y1b, a1b, f1b, y2b, pb, pdb = zeros(N, EH), zeros(N, EH), zeros(N, EH), zeros(N, E0), zeros(N, E0), zeros(N, E0)
    scratch = zeros(int(nf))
    ncta = C.c_int(0)
    ok(L.t4k_linear_act_head_train(layer, ptr(Xd), ptr(W1d), ptr(B1d), ptr(y1b), ptr(a1b), ptr(f1b), 0.1, ptr(W2d), ptr(B2d), ptr(y2b), ptr(pb), ptr(pdb),
                                   ptr(Td), ptr(scratch), C.byref(ncta), N, EH, E1, E0, None), "train tail")
    dwb, dbb, db1b = dev(dW20), dev(dB20), dev(dB10)
    ok(L.t4k_head_grad_finish(ptr(scratch), ncta.value, E0, EH, ptr(dwb), ptr(dbb), ptr(db1b), None))
    for a, b, nm in ((y1a, y1b, "Y1 <- dY1"), (a1a, a1b, "A1 <- dX2"), (f1a, f1b, "F1"), (y2a, y2b, "Ylin <- p - y"), (pa, pb, "P <- p - y"), (pda, pdb, "Pdup = p")):
        assert_exact(host(b), host(a), nm)
    assert_close(host(dwb), host(dwa), rtol=1e-5, what="dW2"); assert_close(host(dbb), host(dba), rtol=1e-5, what="dB2")
    assert_close(host(db1b), host(db1a), rtol=1e-5, what="dB1")


@pytest.mark.parametrize("N,E0,E1", [(512, 100, 1960), (1024, 512, 784), (1024, 256, 512), (96, 36, 48)])
def test_linear_bwd_pair(N, E0, E1):
    """dX and dW of a linear layer in one launch of the layer GEMM == the two GEMM calls (FP32-grade; splits may differ)"""
    X, W, dY, dW0 = rnd(N, E1), rnd(E0, E1), rnd(N, E0), rnd(E0, E1)
    dx, dw = zeros(N, E1), dev(dW0)
    rc = lib().t4k_linear_bwd_pair(ptr(dev(X)), ptr(dev(W)), ptr(dev(dY)), ptr(dx), ptr(dw), N, E0, E1, None)
    if rc == t4.ENOSUP:
        pytest.skip("pair does not fit one co-resident wave: the caller uses t4k_linear_bwd_ex")
    ok(rc, "bwd pair")
    assert_close(host(dx), orc.gemm(dY, W), rtol=1e-4, what="dX")
    assert_close(host(dw), orc.gemm(dY, X, O=dW0, alpha=1.0, beta=1.0, tA=True), rtol=1e-4, what="dW")
    assert_close(host(dx), gemm_ref64(dY, W, np.zeros((N, E1)), 1.0, 0.0, 0, 0), rtol=1e-5, what="dX vs f64")
    assert_close(host(dw), gemm_ref64(dY, X, dW0, 1.0, 1.0, 1, 0), rtol=1e-5, what="dW vs f64")


@pytest.mark.parametrize("N,E0,E1", [(1024, 512, 784), (512, 100, 1960), (1024, 256, 512), (96, 36, 48)])
def test_linear_bwd_act(N, E0, E1):
    """_blinear + the _bactivate in front of it, mask multiply in the dX GEMM's epilogue: same tensors as the two calls"""
    X, W, dY = rnd(N, E1), rnd(E0, E1), rnd(N, E0)
    F = (rnd(N, E1) > 0).astype(np.float32) * 0.8 + 0.2
    dW0, dB0 = rnd(E0, E1), rnd(E0)
    dx, dxp, dw, db = zeros(N, E1), zeros(N, E1), dev(dW0), dev(dB0)
    ok(lib().t4k_linear_bwd_act(ptr(dev(X)), ptr(dev(W)), ptr(dev(dY)), ptr(dx), ptr(dw), ptr(db), ptr(dev(F)), ptr(dxp), N, E0, E1, 1, 0, None))
    rdb = dB0.copy(); orc.lib().orc_dlinear_db(orc._p(dY), orc._p(rdb), N, E0)
    assert_close(host(db), rdb, rtol=1e-4, what="dB")
    assert_close(host(dw), orc.gemm(dY, X, O=dW0, alpha=1.0, beta=1.0, tA=True), rtol=1e-4, what="dW")
    assert_close(host(dx), orc.gemm(dY, W), rtol=1e-4, what="dX")
    assert_exact(host(dxp), orc.tt_op(orc.MUL, host(dx), F), "dX * F")
    # in place (X and dX share a buffer, as Model::_blinear calls it)
    xio, dw2, db2 = dev(X), dev(dW0), dev(dB0)
    ok(lib().t4k_linear_bwd_act(ptr(xio), ptr(dev(W)), ptr(dev(dY)), ptr(xio), ptr(dw2), ptr(db2), ptr(dev(F)), ptr(dxp), N, E0, E1, 1, 0, None))
    assert_exact(host(xio), host(dx)); assert_exact(host(dw2), host(dw))


@pytest.mark.parametrize("engine", [t4.GEMM_TC, t4.GEMM_TC_BF16X3])
@pytest.mark.parametrize("tA,tB,M,N,K", [(0, 0, 2200, 2400, 1024), (1, 1, 2304, 2100, 1100), (0, 1, 4096, 2048, 2048)])
def test_gemm_tc_cta_pairs_ragged(engine, tA, tB, M, N, K):
    """the CTA-pair kernel (k_gemm_tc2: cta_group::2, 256 x 256 tiles, three stages) on ragged M / N / K, transposed operands, alpha / beta,
    with and without a K-split tail wave: FP32-grade against float64"""
    g = torch.Generator(device="cuda").manual_seed(7)
    A = torch.rand((K, M) if tA else (M, K), device="cuda", generator=g) * 2 - 1
    B = torch.rand((N, K) if tB else (K, N), device="cuda", generator=g) * 2 - 1
    O0 = torch.rand(M, N, device="cuda", generator=g)
    o = O0.clone()
    ok(lib().t4k_gemm_ex(engine, ptr(A), ptr(B), ptr(o), 0.5, 2.0, tA, tB, M, N, K, 1, 1, 0, 0, 0, None), "pair gemm")
    idx = torch.cat([torch.arange(0, M, 97, device="cuda"), torch.tensor([M - 1], device="cuda")])
    a = (A.t() if tA else A)[idx].double(); b = (B.t() if tB else B).double()
    ref = 0.5 * (a @ b) + 2.0 * O0[idx].double()
    got = o[idx].double()
    rms = float((got - ref).pow(2).mean().sqrt() / ref.pow(2).mean().sqrt())
    assert rms < (1.2e-5 if engine == t4.GEMM_TC_BF16X3 else 4e-6), rms
    assert float((got - ref).abs().max() / ref.abs().max()) < 1e-4
    o2 = O0.clone()                                               # deterministic
    ok(lib().t4k_gemm_ex(engine, ptr(A), ptr(B), ptr(o2), 0.5, 2.0, tA, tB, M, N, K, 1, 1, 0, 0, 0, None))
    assert torch.equal(o, o2)


@pytest.mark.parametrize("tA,tB", [(0, 0), (1, 1)])
def test_gemm_tc_tail_split_ragged(tA, tB):
    """170 tiles on 148 SMs: the 22 tiles of the last wave are cut into K slices (gemm_tc.cu tail split); ragged M and N, alpha/beta"""
    M, N, K = 2100, 2400, 1024
    g = torch.Generator(device="cuda").manual_seed(3)
    A = torch.rand((K, M) if tA else (M, K), device="cuda", generator=g) * 2 - 1
    B = torch.rand((N, K) if tB else (K, N), device="cuda", generator=g) * 2 - 1
    O0 = torch.rand(M, N, device="cuda", generator=g) * 2 - 1
    o = O0.clone()
    ok(lib().t4k_gemm_ex(t4.GEMM_TC, ptr(A), ptr(B), ptr(o), 0.5, 2.0, tA, tB, M, N, K, 1, 1, 0, 0, 0, None), "gemm_tc tail")
    a = (A.double().T if tA else A.double()); b = (B.double().T if tB else B.double())
    ref = (0.5 * (a @ b) + 2.0 * O0.double()).cpu().numpy()
    assert_close(o.cpu().numpy(), ref, rtol=1e-5, what="tail split vs f64")


@pytest.mark.parametrize("tA,tB,M,N,K", [(0, 0, 1024, 1024, 1024), (1, 0, 300, 520, 777), (0, 1, 2100, 2400, 1024), (1, 1, 129, 257, 36)])
def test_gemm_bf16x3_engine(tA, tB, M, N, K):
    """opt-in BF16x3 engine (a = hi + lo in bf16, three MMAs at twice the TF32 rate): ~1e-5 of the result's rms — inside the
    north star's 1e-4 with 10x margin, but 5x coarser than 3xTF32, which stays the default"""
    A = rnd(K, M) if tA else rnd(M, K)
    B = rnd(N, K) if tB else rnd(K, N)
    O0 = rnd(M, N)
    o = dev(O0)
    ok(lib().t4k_gemm_ex(t4.GEMM_TC_BF16X3, ptr(dev(A)), ptr(dev(B)), ptr(o), 0.5, 2.0, tA, tB, M, N, K, 1, 1, 0, 0, 0, None), "gemm bf16x3")
    ref = gemm_ref64(A, B, O0, 0.5, 2.0, tA, tB)
    got = host(o)
    assert_close(got, ref, rtol=3e-5, what="bf16x3 vs f64")
    rms = np.sqrt(np.mean((got - ref) ** 2)) / np.sqrt(np.mean(ref ** 2))
    assert rms < 1e-5, rms


def test_gemm_4096_property():
    # BASELINE size (config 2): size-independent checks — linearity in alpha and the row-sum identity
    # (A@B)·1 = A·(B·1), evaluated in float64 on the host in O(n^2).
    n = 4096
    g = torch.Generator(device="cuda").manual_seed(7)
    A = torch.rand(n, n, device="cuda", generator=g) * 2 - 1
    B = torch.rand(n, n, device="cuda", generator=g) * 2 - 1
    o1, o2 = zeros(n, n), zeros(n, n)
    ok(lib().t4k_gemm(ptr(A), ptr(B), ptr(o1), 1.0, 0.0, 0, 0, n, n, n, 1, 1, 0, 0, 0, None))
    ok(lib().t4k_gemm(ptr(A), ptr(B), ptr(o2), 2.0, 0.0, 0, 0, n, n, n, 1, 1, 0, 0, 0, None))
    assert torch.equal(o2, 2 * o1)                        # exact: scaling by 2 commutes with rounding
    rs = o1.double().sum(1).cpu().numpy()
    ref = (A.double() @ B.double().sum(1)).cpu().numpy()
    # a row sum adds 4096 elements whose FP32-grade errors (~2e-6 rel, tensor-core accumulation
    # truncates toward zero so they do not cancel) → 1e-4 of the largest row sum
    assert_close(rs, ref, rtol=1e-4, atol=1e-4 * np.abs(ref).max())
    # spot-check 64 full rows against float64
    idx = torch.arange(0, n, 64, device="cuda")
    ref_rows = (A[idx].double() @ B.double()).cpu().numpy()
    # AUTO routes this size class to the BF16x3 engine: rms error ~4e-6 of the result (the reference's own FP32-FMA accumulation
    # sits at ~sqrt(K) 2^-24 = 3.8e-6); the elementwise bar is the north star's 1e-4 with a 3x margin, the rms bar 1e-5
    got_rows = o1[idx].cpu().numpy()
    assert_close(got_rows, ref_rows, rtol=3e-5, what="4096 rows")
    assert np.sqrt(np.mean((got_rows - ref_rows) ** 2)) / np.sqrt(np.mean(ref_rows ** 2)) < 1e-5
    # the 3xTF32 engine on the same size keeps the tighter bar
    o3 = zeros(n, n)
    ok(lib().t4k_gemm_ex(t4.GEMM_TC, ptr(A), ptr(B), ptr(o3), 1.0, 0.0, 0, 0, n, n, n, 1, 1, 0, 0, 0, None))
    assert_close(o3[idx].cpu().numpy(), ref_rows, rtol=1e-5, what="4096 rows, 3xTF32")


@pytest.mark.parametrize("engine", [t4.GEMM_TC, t4.GEMM_AUTO])
def test_gemm_4096_vs_reference_kernel(engine):
    """BASELINE config 2 at full size against the REFERENCE'S OWN kernel (k_gemm_tile_claude, src/t4math.cu:478-583, run through
    oracle/_ref/refkern on this GPU): both tensor-core engines (3xTF32, and BF16x3 which AUTO selects at this size) within the north
    star's 1e-4 of the reference output"""
    from oracle import refkern
    if not refkern.available():
        pytest.skip("oracle/_ref/refkern not built (needs the reference sources; built in the build container)")
    n = 4096
    rng = np.random.default_rng(5)
    A = (rng.random((n, n), dtype=np.float32) * 2 - 1).astype(np.float32)
    B = (rng.random((n, n), dtype=np.float32) * 2 - 1).astype(np.float32)
    ref = refkern.one("gemm", ints=(3, 0, 0, n, n, n, 1), flts=(1.0, 0.0), arrs=(A, B, np.zeros((n, n), np.float32)))[0].reshape(n, n)
    o = zeros(n, n)
    ok(lib().t4k_gemm_ex(engine, ptr(dev(A)), ptr(dev(B)), ptr(o), 1.0, 0.0, 0, 0, n, n, n, 1, 1, 0, 0, 0, None), "gemm 4096")
    got = host(o)
    assert_close(got, ref, rtol=1e-4, what="4096^3 vs the reference kernel")
    rms = np.sqrt(np.mean((got.astype(np.float64) - ref) ** 2)) / np.sqrt(np.mean(ref.astype(np.float64) ** 2))
    assert rms < 1e-5, rms


# ------------------------------------------------------------------ fused linear epilogues / classifier head
@pytest.mark.parametrize("layer,alpha", [(t4.L_RELU, 0.0), (t4.L_LEAKYRL, 0.2), (t4.L_TANH, 0.0), (t4.L_SIGMOID, 0.0), (t4.L_ELU, 1.0), (t4.L_SELU, 0.0)])
@pytest.mark.parametrize("N,E0,E1", [(512, 100, 1960), (3, 5, 7), (1024, 512, 784)])
def test_linear_act_fwd(layer, alpha, N, E0, E1):
    X, W, Bv = rnd(N, E1), rnd(E0, E1) * 0.1, rnd(E0)
    y, a, f = zeros(N, E0), zeros(N, E0), zeros(N, E0)
    ok(lib().t4k_linear_act_fwd(layer, ptr(dev(X)), ptr(dev(W)), ptr(dev(Bv)), ptr(y), ptr(a), ptr(f), alpha, N, E0, E1, None))
    ref = orc.gemm(X, W, tB=True); orc.lib().orc_bias(orc._p(Bv), orc._p(ref), N, E0)
    assert_close(host(y), ref, rtol=1e-4, what="linear out")
    ra, rf = orc.activate(layer, host(y), alpha)            # activation of OUR pre-activation: isolates the epilogue
    assert_close(host(a), ra, rtol=1e-5, what="act out"); assert_close(host(f), rf, rtol=1e-5, what="act mask")


@pytest.mark.parametrize("N,E0,E1", [(512, 10, 100), (7, 3, 5), (64, 32, 128), (33, 1, 256)])
def test_mlp_head_fwd(N, E0, E1):
    X, W, Bv = rnd(N, E1), rnd(E0, E1), rnd(E0)
    y, pr = zeros(N, E0), zeros(N, E0)
    ok(lib().t4k_mlp_head_fwd(ptr(dev(X)), ptr(dev(W)), ptr(dev(Bv)), ptr(y), ptr(pr), N, E0, E1, None))
    ref = orc.gemm(X, W, tB=True); orc.lib().orc_bias(orc._p(Bv), orc._p(ref), N, E0)
    assert_close(host(y), ref, rtol=1e-4, what="head linear")
    assert_close(host(pr), orc.softmax(ref, N), rtol=1e-4, what="head softmax")
    assert lib().t4k_mlp_head_fwd(ptr(dev(X)), ptr(dev(W)), ptr(dev(Bv)), ptr(y), ptr(pr), N, 33, E1, None) == t4.ENOSUP


@pytest.mark.parametrize("N,E0,E1,act,prev", [(512, 10, 100, True, True), (5, 3, 7, True, False), (64, 32, 64, False, True), (40, 16, 128, True, True), (9, 2, 33, False, False)])
def test_mlp_head_bwd(N, E0, E1, act, prev):
    P, T, X2, W = np.abs(rnd(N, E0)), orc.onehot(np.arange(N) % E0, E0), rnd(N, E1), rnd(E0, E1)
    F1 = (rnd(N, E1) > 0).astype(np.float32)
    dW0, dB0, dB10 = rnd(E0, E1), rnd(E0), rnd(E1)
    p, yl, x2, y1 = dev(P), zeros(N, E0), dev(X2), zeros(N, E1)
    dw, db, db1 = dev(dW0), dev(dB0), dev(dB10)
    for rep in range(2):                                    # twice: the arrival counter must re-arm itself
        p, x2, dw, db, db1 = dev(P), dev(X2), dev(dW0), dev(dB0), dev(dB10)
        ok(lib().t4k_mlp_head_bwd(ptr(p), ptr(dev(T)), ptr(yl), ptr(x2), ptr(dev(F1)) if act else None, ptr(y1) if act else None, ptr(dev(W)),
                                  ptr(dw), ptr(db), ptr(db1) if prev else None, N, E0, E1, 1, None))
        d = orc.tt_op(orc.SUB, P, T)
        assert_exact(host(p), d); assert_exact(host(yl), d)
        dx = orc.gemm(d, W)
        assert_close(host(x2), dx, rtol=1e-4, what="head dX")
        rdb = dB0.copy(); orc.lib().orc_dlinear_db(orc._p(d), orc._p(rdb), N, E0)
        assert_close(host(db), rdb, rtol=1e-4, what="head dB")
        assert_close(host(dw), orc.gemm(d, X2, O=dW0, alpha=1.0, beta=1.0, tA=True), rtol=1e-4, what="head dW")
        g = host(x2) * F1 if act else host(x2)
        if act:
            assert_exact(host(y1), g)
        if prev:
            rdb1 = dB10.copy(); orc.lib().orc_dlinear_db(orc._p(np.ascontiguousarray(g)), orc._p(rdb1), N, E1)
            assert_close(host(db1), rdb1, rtol=1e-4, what="dB of the linear in front")
    # train == 0: parameter gradients untouched
    p, x2, dw, db = dev(P), dev(X2), dev(dW0), dev(dB0)
    ok(lib().t4k_mlp_head_bwd(ptr(p), ptr(dev(T)), ptr(yl), ptr(x2), None, None, ptr(dev(W)), ptr(dw), ptr(db), None, N, E0, E1, 0, None))
    assert_exact(host(dw), dW0); assert_exact(host(db), dB0)
    assert lib().t4k_mlp_head_bwd(ptr(p), ptr(dev(T)), ptr(yl), ptr(x2), None, None, ptr(dev(W)), ptr(dw), ptr(db), None, N, E0, 129, 0, None) == t4.ENOSUP


# ------------------------------------------------------------------ linear / activation / softmax
@pytest.mark.parametrize("N,E0,E1", [(1, 3, 2), (3, 2, 2), (512, 100, 1960), (512, 10, 100), (1024, 512, 784)])
def test_linear_fwd_bwd(N, E0, E1):
    X, W, Bv, dY = rnd(N, E1), rnd(E0, E1), rnd(E0), rnd(N, E0)
    y = zeros(N, E0)
    ok(lib().t4k_linear_fwd(ptr(dev(X)), ptr(dev(W)), ptr(dev(Bv)), ptr(y), N, E0, E1, None))
    ref = orc.gemm(X, W, tB=True); orc.lib().orc_bias(orc._p(Bv), orc._p(ref), N, E0)
    assert_close(host(y), ref, rtol=1e-4, what="linear fwd")
    dW0, dB0 = rnd(E0, E1), rnd(E0)
    dx, dw, db = zeros(N, E1), dev(dW0), dev(dB0)
    ok(lib().t4k_linear_bwd(ptr(dev(X)), ptr(dev(W)), ptr(dev(dY)), ptr(dx), ptr(dw), ptr(db), N, E0, E1, 1, None))
    rdb = dB0.copy(); orc.lib().orc_dlinear_db(orc._p(dY), orc._p(rdb), N, E0)
    assert_close(host(db), rdb, rtol=1e-4, what="dB")
    assert_close(host(dw), orc.gemm(dY, X, O=dW0, alpha=1.0, beta=1.0, tA=True), rtol=1e-4, what="dW")
    assert_close(host(dx), orc.gemm(dY, W), rtol=1e-4, what="dX")
    # in place (dX stored over X, as Model::_blinear calls it): with the one-launch dX/dW pair the dX tiles are held back until the dW problem has
    # read X — repeated, a lost ordering would show as a dW built from overwritten rows
    for _ in range(4):
        xio, dw3, db3 = dev(X), dev(dW0), dev(dB0)
        ok(lib().t4k_linear_bwd(ptr(xio), ptr(dev(W)), ptr(dev(dY)), ptr(xio), ptr(dw3), ptr(db3), N, E0, E1, 1, None))
        assert_exact(host(xio), host(dx), "dX in place"); assert_exact(host(dw3), host(dw), "dW with dX in place")
    # train == 0: parameters' gradients untouched
    dw2, db2 = dev(dW0), dev(dB0)
    ok(lib().t4k_linear_bwd(ptr(dev(X)), ptr(dev(W)), ptr(dev(dY)), ptr(dx), ptr(dw2), ptr(db2), N, E0, E1, 0, None))
    assert_exact(host(dw2), dW0); assert_exact(host(db2), dB0)


@pytest.mark.parametrize("n,rate", [(7, 0.3), (4096, 0.5), (100352, 0.3), (1024 * 512, 0.3)])
def test_dropout_fwd_equals_rand_then_activate(n, rate):
    """the one-launch dropout forward draws the mask t4k_rand would draw and applies k_activate's L_DROPOUT rule: same bits, same stream position"""
    L = lib()
    x = dev(rnd(n))
    L.t4k_rand_seed(123)
    f1, o1, nx1 = zeros(n), zeros(n), zeros(64)
    ok(L.t4k_rand(ptr(f1), n, t4.UNIFORM, 0.0, 1.0, None)); ok(L.t4k_activate_fwd(t4.L_DROPOUT, ptr(x), ptr(o1), ptr(f1), rate, n, None))
    ok(L.t4k_rand(ptr(nx1), 64, t4.UNIFORM, 0.0, 1.0, None))
    L.t4k_rand_seed(123)
    f2, o2, nx2 = zeros(n), zeros(n), zeros(64)
    ok(L.t4k_dropout_fwd(ptr(x), ptr(o2), ptr(f2), rate, n, 0, n, None)); ok(L.t4k_rand(ptr(nx2), 64, t4.UNIFORM, 0.0, 1.0, None))
    assert_exact(host(f2), host(f1), "mask"); assert_exact(host(o2), host(o1), "output"); assert_exact(host(nx2), host(nx1), "next draw")
    kept = float(host(f2).mean())
    assert abs(kept - (1.0 - rate)) < 5.0 / np.sqrt(n) + 1e-9 or n < 100
    # a shard of a larger batch-major tensor: the slice of the single-device mask
    if n % 4 == 0 and n >= 4096:
        L.t4k_rand_seed(123)
        q = n // 4
        f3, o3 = zeros(q), zeros(q)
        xs = x[2 * q:3 * q].clone()
        ok(L.t4k_dropout_fwd(ptr(xs), ptr(o3), ptr(f3), rate, q, 2 * q, n, None))
        assert_exact(host(f3), host(f1)[2 * q:3 * q], "shard mask"); assert_exact(host(o3), host(o1)[2 * q:3 * q], "shard output")


ACTS = [(t4.L_RELU, 0.0), (t4.L_TANH, 0.0), (t4.L_SIGMOID, 0.0), (t4.L_SELU, 0.0), (t4.L_LEAKYRL, 0.2),
        (t4.L_ELU, 1.0), (t4.L_DROPOUT, 0.3)]


@pytest.mark.parametrize("layer,alpha", ACTS)
@pytest.mark.parametrize("n", [7, 4096, 100352])
def test_activate(layer, alpha, n):
    x = rnd(n, lo=-3, hi=3)
    x[:3] = [0.0, -0.0, 1e-30]
    u = RNG.random(n, dtype=np.float32)
    o, f = zeros(n), dev(u)
    ok(lib().t4k_activate_fwd(layer, ptr(dev(x)), ptr(o), ptr(f), alpha, n, None))
    ro, rf = orc.activate(layer, x, alpha, mask=u)
    if layer in (t4.L_RELU, t4.L_LEAKYRL, t4.L_DROPOUT):
        assert_exact(host(o), ro); assert_exact(host(f), rf)               # routing / mask: bit exact
    else:
        assert_close(host(o), ro, rtol=1e-5, atol=1e-6); assert_close(host(f), rf, rtol=1e-5, atol=1e-6)
    dy, dx = rnd(n), zeros(n)
    ok(lib().t4k_activate_bwd(ptr(dev(dy)), ptr(f), ptr(dx), n, None))
    assert_exact(host(dx), orc.tt_op(orc.MUL, dy, host(f)))


@pytest.mark.parametrize("N,Cc", [(1, 2), (512, 10), (33, 300), (4, 1000)])
def test_softmax_logsoftmax(N, Cc):
    x = rnd(N, Cc, lo=-4, hi=4)
    o = zeros(N, Cc)
    ok(lib().t4k_softmax_fwd(ptr(dev(x)), ptr(o), N, Cc, None))
    assert_close(host(o), orc.softmax(x, N), rtol=1e-5, atol=1e-7)
    ok(lib().t4k_logsoftmax_fwd(ptr(dev(x)), ptr(o), N, Cc, None))
    assert_close(host(o), orc.logsoftmax(x, N), rtol=1e-5, atol=1e-5)


# ------------------------------------------------------------------ conv / pool / batchnorm
CONV_CASES = [  # N,H,W,C1,C0,K,S,P
    (2, 16, 16, 1, 2, 3, 1, 1),          # t4_30d.4th toy CNN, small-channel path
    (4, 28, 28, 1, 10, 3, 1, 1),         # MNIST first layer (config 3)
    (3, 12, 12, 3, 5, 5, 1, 2),
    (2, 9, 9, 2, 4, 1, 1, 0),
    (2, 14, 14, 2, 6, 4, 2, 1),          # the 4x4 stride-2 config
    (2, 14, 14, 16, 32, 3, 1, 1),        # general implicit-GEMM path
    (1, 10, 10, 20, 70, 5, 1, 2),
    (2, 8, 8, 24, 24, 4, 2, 1),
    (1, 56, 56, 64, 64, 3, 1, 1),        # one sample of config 5
]


@pytest.mark.parametrize("N,H1,W1,C1,C0", [(2, 7, 7, 6, 2), (3, 4, 4, 8, 6), (2, 8, 8, 6, 1), (2, 7, 7, 24, 24), (1, 14, 14, 64, 32)])
def test_conv_transpose_fwd_bwd(N, H1, W1, C1, C0):
    """L_DCONV (word `dconv2d`: 4x4, stride 2, padding 1): t4k_dconv2d_fwd / _bwd against the oracle's restatement — the reference's k_dconv2d /
    k_conv2d with swapped roles (forward.cu:110, backprop.cu:137), itself pinned to the reference kernels' outputs (tests/test_golden_ref_kernels.py)"""
    K, S, P = 4, 2, 1
    H0, W0 = orc.convt_out_dims(H1, W1, K, S, P)
    I, F, Bv = rnd(N, H1, W1, C1), rnd(C0, K, K, C1, lo=-.3, hi=.3), rnd(C0)
    o = dev(np.full((N, H0, W0, C0), np.nan, np.float32))                    # must not need pre-zeroing
    ok(lib().t4k_dconv2d_fwd(ptr(dev(I)), ptr(dev(F)), ptr(dev(Bv)), ptr(o), N, H1, W1, C1, H0, W0, C0, K, S, P, None), "dconv fwd")
    assert_close(host(o), orc.convt2d(I, F, Bv, K, S, P), rtol=1e-4, what="conv-transpose fwd")
    dO, dF0, dB0 = rnd(N, H0, W0, C0), rnd(C0, K, K, C1), rnd(C0)
    dx, df, db = dev(np.full(I.shape, np.nan, np.float32)), dev(dF0), dev(dB0)
    ok(lib().t4k_dconv2d_bwd(ptr(dev(I)), ptr(dev(dO)), ptr(dev(F)), ptr(dx), ptr(df), ptr(db), N, H1, W1, C1, H0, W0, C0, K, S, P, 1, None), "dconv bwd")
    rdx, rdf, rdb = orc.dconvt2d(I, dO, F, K, S, P, dF0, dB0, True)
    assert_close(host(dx), rdx, rtol=1e-4, what="conv-transpose dX")
    assert_close(host(df), rdf, rtol=1e-4, what="conv-transpose dF")
    assert_close(host(db), rdb, rtol=1e-4, what="conv-transpose dB")
    # not trainable: parameters untouched
    df2, db2 = dev(dF0), dev(dB0)
    ok(lib().t4k_dconv2d_bwd(ptr(dev(I)), ptr(dev(dO)), ptr(dev(F)), ptr(dx), ptr(df2), ptr(db2), N, H1, W1, C1, H0, W0, C0, K, S, P, 0, None))
    assert_exact(host(df2), dF0, "dF untouched"); assert_exact(host(db2), dB0, "dB untouched")
    # geometry the layer cannot have / aliasing
    assert lib().t4k_dconv2d_fwd(ptr(dev(I)), ptr(dev(F)), ptr(dev(Bv)), ptr(o), N, H1, W1, C1, H0 + 2, W0, C0, K, S, P, None) == t4.EINVAL
    assert lib().t4k_dconv2d_fwd(ptr(dev(I)), ptr(dev(F)), ptr(dev(Bv)), ptr(o), N, H1, W1, C1, H0, W0, C0, 2, 2, 0, None) == t4.ENOSUP


@pytest.mark.parametrize("N,H,W,C1,C0,K,S,P", CONV_CASES)
def test_conv2d_fwd_bwd(N, H, W, C1, C0, K, S, P):
    I, F, Bv = rnd(N, H, W, C1), rnd(C1, K, K, C0, lo=-.3, hi=.3), rnd(C0)
    H0, W0 = orc.conv_out_dims(H, W, K, S, P)
    o = dev(np.full((N, H0, W0, C0), np.nan, np.float32))                    # must not need pre-zeroing
    ok(lib().t4k_conv2d_fwd(ptr(dev(I)), ptr(dev(F)), ptr(dev(Bv)), ptr(o), N, H, W, C1, H0, W0, C0, K, S, P, None))
    assert_close(host(o), orc.conv2d(I, F, Bv, K, S, P), rtol=1e-4, what="conv fwd")
    dO = rnd(N, H0, W0, C0)
    dF0, dB0 = rnd(C1, K, K, C0), rnd(C0)
    dx, df, db = dev(np.full(I.shape, np.nan, np.float32)), dev(dF0), dev(dB0)
    ok(lib().t4k_conv2d_bwd(ptr(dev(I)), ptr(dev(dO)), ptr(dev(F)), ptr(dx), ptr(df), ptr(db),
                            N, H, W, C1, H0, W0, C0, K, S, P, 1, None))
    rdx, rdf, rdb = orc.dconv2d(I, dO, F, K, S, P, dF0, dB0, True)
    assert_close(host(dx), rdx, rtol=1e-4, what="conv dX (flipped taps)")
    assert_close(host(df), rdf, rtol=1e-4, what="conv dF")
    assert_close(host(db), rdb, rtol=1e-4, what="conv dB")
    df2, db2 = dev(dF0), dev(dB0)
    ok(lib().t4k_conv2d_bwd(ptr(dev(I)), ptr(dev(dO)), ptr(dev(F)), ptr(dx), ptr(df2), ptr(db2),
                            N, H, W, C1, H0, W0, C0, K, S, P, 0, None))
    assert_exact(host(df2), dF0); assert_exact(host(db2), dB0)               # train == 0
    assert lib().t4k_conv2d_fwd(ptr(dev(I)), ptr(dev(F)), ptr(dev(Bv)), ptr(o), N, H, W, C1, H0, W0, C0, 7, 1, 3, None) == -2


CONV_TC_CASES = [  # N,H,W,C1,C0,K,P  (stride 1, "same"): tcgen05 implicit-GEMM engine
    (2, 12, 12, 32, 16, 3, 1),           # CK=32, smallest N
    (3, 9, 9, 64, 64, 3, 1),             # 243 pixels: ragged last tile
    (1, 16, 16, 64, 128, 5, 2),          # 5x5, C0=128 → single accumulator buffer
    (2, 10, 10, 96, 48, 1, 0),           # 1x1, three 32-channel chunks
    (3, 12, 12, 32, 64, 3, 1),           # wgrad: 4 taps per M-tile
    (2, 20, 20, 64, 32, 1, 0),           # wgrad 1x1
    (9, 16, 16, 64, 64, 3, 1),           # wgrad: 36 k-blocks
    (5, 28, 28, 64, 32, 3, 1),           # 31 tiles → several per CTA? (no: one each) exercises tile loop bounds
    (40, 56, 56, 64, 64, 3, 1),          # 980 tiles on 148 CTAs: persistent loop, double-buffered accumulators
]


@pytest.mark.parametrize("N,H,W,C1,C0,K,P", CONV_TC_CASES)
def test_conv2d_tc_engine(N, H, W, C1, C0, K, P):
    """fwd + dX on the tensor-core engine: vs the oracle where it finishes in seconds, vs the CUDA-core engine always"""
    I, F, Bv, dO = rnd(N, H, W, C1), rnd(C1, K, K, C0, lo=-.3, hi=.3), rnd(C0), rnd(N, H, W, C0)
    Id, Fd, Bd, dOd = dev(I), dev(F), dev(Bv), dev(dO)
    o_tc, o_si = dev(np.full((N, H, W, C0), np.nan, np.float32)), zeros(N, H, W, C0)
    dx_tc, dx_si = dev(np.full((N, H, W, C1), np.nan, np.float32)), zeros(N, H, W, C1)
    try:
        ok(lib().t4k_set_conv_engine(t4.GEMM_TC))
        ok(lib().t4k_conv2d_fwd(ptr(Id), ptr(Fd), ptr(Bd), ptr(o_tc), N, H, W, C1, H, W, C0, K, 1, P, None), "conv fwd tc")
        ok(lib().t4k_conv2d_bwd(ptr(Id), ptr(dOd), ptr(Fd), ptr(dx_tc), None, None, N, H, W, C1, H, W, C0, K, 1, P, 0, None), "conv dX tc")
        ok(lib().t4k_set_conv_engine(t4.GEMM_SIMT))
        ok(lib().t4k_conv2d_fwd(ptr(Id), ptr(Fd), ptr(Bd), ptr(o_si), N, H, W, C1, H, W, C0, K, 1, P, None))
        ok(lib().t4k_conv2d_bwd(ptr(Id), ptr(dOd), ptr(Fd), ptr(dx_si), None, None, N, H, W, C1, H, W, C0, K, 1, P, 0, None))
    finally:
        lib().t4k_set_conv_engine(t4.GEMM_AUTO)
    if (C1 in (32, 64)) and (C0 in (32, 64)) and K in (1, 3):                 # tensor-core weight gradient
        dF0, dB0 = rnd(C1, K, K, C0), rnd(C0)
        df_tc, db_tc, df_si, db_si = dev(dF0), dev(dB0), dev(dF0), dev(dB0)
        try:
            ok(lib().t4k_set_conv_engine(t4.GEMM_TC))
            ok(lib().t4k_conv2d_bwd(ptr(Id), ptr(dOd), ptr(Fd), ptr(dx_tc), ptr(df_tc), ptr(db_tc), N, H, W, C1, H, W, C0, K, 1, P, 1, None), "conv bwd tc")
            ok(lib().t4k_set_conv_engine(t4.GEMM_SIMT))
            ok(lib().t4k_conv2d_bwd(ptr(Id), ptr(dOd), ptr(Fd), ptr(dx_si), ptr(df_si), ptr(db_si), N, H, W, C1, H, W, C0, K, 1, P, 1, None))
        finally:
            lib().t4k_set_conv_engine(t4.GEMM_AUTO)
        assert_close(host(df_tc), host(df_si), rtol=2e-5, what="conv dF tc vs simt")
        assert_close(host(db_tc), host(db_si), rtol=2e-5, what="conv dB tc vs simt")
        if N * H * W <= 4096:
            _, rdf, rdb = orc.dconv2d(I, dO, F, K, 1, P, dF0, dB0, True)
            assert_close(host(df_tc), rdf, rtol=1e-4, what="conv dF tc vs oracle")
            assert_close(host(db_tc), rdb, rtol=1e-4, what="conv dB tc vs oracle")
    assert_close(host(o_tc), host(o_si), rtol=2e-5, what="conv fwd tc vs simt")
    assert_close(host(dx_tc), host(dx_si), rtol=2e-5, what="conv dX tc vs simt")
    if N * H * W <= 4096:
        assert_close(host(o_tc), orc.conv2d(I, F, Bv, K, 1, P), rtol=1e-4, what="conv fwd tc vs oracle")
        rdx, _, _ = orc.dconv2d(I, dO, F, K, 1, P, np.zeros_like(F), np.zeros_like(Bv), False)
        assert_close(host(dx_tc), rdx, rtol=1e-4, what="conv dX tc vs oracle")


@pytest.mark.parametrize("layer", [t4.L_MAXPOOL, t4.L_AVGPOOL, t4.L_MINPOOL, t4.L_USAMPLE])
@pytest.mark.parametrize("N,H,W,Cc,K", [(2, 4, 4, 1, 2), (512, 28, 28, 10, 2), (3, 9, 12, 5, 3), (2, 6, 6, 33, 2)])
def test_pool_fwd_bwd(layer, N, H, W, Cc, K):
    x = rnd(N, H, W, Cc)
    x[0, :K, :K, 0] = 0.5                                                    # a full tie: first element must win
    o = zeros(N, H // K, W // K, Cc)
    ok(lib().t4k_pool_fwd(layer, ptr(dev(x)), ptr(o), N, H, W, H // K, W // K, Cc, K, None))
    ref = orc.pool(layer, x, K)
    (assert_exact if layer in (t4.L_MAXPOOL, t4.L_MINPOOL) else assert_close)(host(o), ref)
    dy = rnd(N, H // K, W // K, Cc)
    xi = dev(x)
    ok(lib().t4k_pool_bwd(layer, ptr(xi), ptr(dev(dy)), N, H, W, H // K, W // K, Cc, K, None))
    rb = orc.dpool(layer, x, dy, K)
    (assert_close if layer == t4.L_AVGPOOL else assert_exact)(host(xi), rb)  # routing is exact


@pytest.mark.parametrize("N,HW,Cc", [(2, 16, 3), (8, 49, 10), (4, 100, 64), (3, 7, 300)])
def test_batchnorm_fwd_bwd(N, HW, Cc):
    x = rnd(N, HW, Cc, lo=-2, hi=3)
    g, b = rnd(Cc, lo=.5, hi=1.5), rnd(Cc)
    o, xh, scr = zeros(N, HW, Cc), zeros(N, HW, Cc), zeros(3 * Cc)
    ok(lib().t4k_batchnorm_fwd(ptr(dev(x)), ptr(o), ptr(xh), ptr(dev(g)), ptr(dev(b)), ptr(scr), N, HW, Cc, None))
    ro, rxh, ravg, rrvar = orc.batchnorm(x, g, b)
    assert_close(host(o), ro, rtol=1e-4, what="bn out"); assert_close(host(xh), rxh, rtol=1e-4, what="bn xhat")
    assert_close(host(scr)[:Cc], rrvar, rtol=1e-4); assert_close(host(scr)[Cc:2 * Cc], ravg, rtol=1e-4, atol=1e-6)
    dy = rnd(N, HW, Cc)
    dW0, dB0 = rnd(Cc), rnd(Cc)
    dx, dw, db = zeros(N, HW, Cc), dev(dW0), dev(dB0)
    ok(lib().t4k_batchnorm_bwd(ptr(dev(dy)), ptr(xh), ptr(dx), ptr(dev(g)), ptr(dw), ptr(db), ptr(scr), N, HW, Cc, 1, None))
    rdx, rdw, rdb = orc.dbatchnorm(dy, rxh, g, rrvar, dW0, dB0, True)
    assert_close(host(dx), rdx, rtol=1e-4, what="bn dX")
    assert_close(host(dw), rdw, rtol=1e-4, atol=1e-6); assert_close(host(db), rdb, rtol=1e-4, atol=1e-6)


# ------------------------------------------------------------------ optimizers / onehot / hit / rand
@pytest.mark.parametrize("n", [5, 1000, 196000])
def test_optimizers(n):
    g0, dg0, m0, v0 = rnd(n), rnd(n), rnd(n, lo=-.1, hi=.1), rnd(n, lo=0, hi=.1)
    for kind in ("sgd0", "sgdm", "adam", "adamw"):
        g, dg, m, v = dev(g0), dev(dg0), dev(m0), dev(v0)
        rg, rdg, rm, rv = g0.copy(), dg0.copy(), m0.copy(), v0.copy()
        P = orc._p
        if kind == "sgd0":
            ok(lib().t4k_sgd(ptr(g), ptr(dg), ptr(m), 3, 0.5, 0.0, n, None)); orc.lib().orc_sgd(P(rg), P(rdg), P(rm), 3, 0.5, 0.0, n)
        elif kind == "sgdm":
            ok(lib().t4k_sgd(ptr(g), ptr(dg), ptr(m), 1, 0.5, 0.9, n, None)); orc.lib().orc_sgd(P(rg), P(rdg), P(rm), 1, 0.5, 0.9, n)
        elif kind == "adam":
            ok(lib().t4k_adam(ptr(g), ptr(dg), ptr(m), ptr(v), 1e-3, 0.9, 0.999, n, None)); orc.lib().orc_adam(P(rg), P(rdg), P(rm), P(rv), 1e-3, 0.9, 0.999, n)
        else:
            ok(lib().t4k_adamw(ptr(g), ptr(dg), ptr(m), ptr(v), 1e-3, 0.9, 0.999, 0.01, n, None)); orc.lib().orc_adamw(P(rg), P(rdg), P(rm), P(rv), 1e-3, 0.9, 0.999, 0.01, n)
        assert_close(host(g), rg, rtol=1e-5, atol=1e-7, what=kind)
        assert not host(dg).any()                                            # dG zeroed (nmath.cu:434,452)
        assert_close(host(m), rm, rtol=1e-5, atol=1e-8, what=kind + " m")
        if kind.startswith("adam"):
            assert_close(host(v), rv, rtol=1e-5, atol=1e-9, what=kind + " v")


def test_optim_multi_matches_per_tensor():
    lens, nws = [90, 10, 196000, 100, 1000, 10], [1, 1, 1, 1, 1, 1]
    nws[0] = 3
    offs = np.concatenate([[0], np.cumsum([(l + 3) // 4 * 4 for l in lens])]).astype(np.int64)
    total = int(offs[-1])
    g0, dg0 = rnd(total), rnd(total)
    seg = np.zeros(len(lens), dtype=[("off", "<i8"), ("len", "<i8"), ("Nw", "<i4"), ("pad", "<i4")])
    seg["off"], seg["len"], seg["Nw"] = offs[:-1], [(l + 3) // 4 * 4 for l in lens], nws
    dseg = torch.from_numpy(seg.view(np.uint8)).cuda()
    for kind in (0, 1, 2):
        g, dg, m, v = dev(g0), dev(dg0), zeros(total), zeros(total)
        ok(lib().t4k_optim_multi(kind, ptr(g), ptr(dg), ptr(m), ptr(v), C.c_void_p(dseg.data_ptr()), len(lens), total, 0.01, 0.9 if kind else 0.0, 0.999, 0.01, None))
        g2, dg2, m2, v2 = dev(g0), dev(dg0), zeros(total), zeros(total)
        for k in range(len(lens)):
            o, ln = int(seg["off"][k]), int(seg["len"][k])
            if kind == 0: ok(lib().t4k_sgd(ptr(g2, o), ptr(dg2, o), ptr(m2, o), int(seg["Nw"][k]), 0.01, 0.0, ln, None))
            elif kind == 1: ok(lib().t4k_adam(ptr(g2, o), ptr(dg2, o), ptr(m2, o), ptr(v2, o), 0.01, 0.9, 0.999, ln, None))
            else: ok(lib().t4k_adamw(ptr(g2, o), ptr(dg2, o), ptr(m2, o), ptr(v2, o), 0.01, 0.9, 0.999, 0.01, ln, None))
        assert_exact(host(g), host(g2)); assert_exact(host(m), host(m2)); assert_exact(host(v), host(v2))


def test_onehot_hit():
    N, E = 512, 10
    lab = RNG.integers(0, 12, N).astype(np.int32)            # labels >= E map to class 0 (loss.cpp:66)
    hot = zeros(N, E)
    dl = torch.from_numpy(lab).cuda()
    ok(lib().t4k_onehot(C.c_void_p(dl.data_ptr()), ptr(hot), N, E, None))
    rh = orc.onehot(lab, E)
    assert_exact(host(hot), rh)
    out = rnd(N, E)
    out[3, :] = 0.25                                          # tie → first index
    cnt = torch.zeros(1, dtype=torch.int32, device="cuda")
    ok(lib().t4k_hit(ptr(dev(out)), ptr(hot), N, E, C.c_void_p(cnt.data_ptr()), None))
    assert int(cnt.cpu()[0]) == orc.hit(out, rh)


def test_rand_statistics_and_reproducibility():
    n = 1 << 20
    a, b = zeros(n), zeros(n)
    ok(lib().t4k_rand_at(ptr(a), n, t4.UNIFORM, 0.0, 1.0, 42, 0, None))
    ok(lib().t4k_rand_at(ptr(b), n, t4.UNIFORM, 0.0, 1.0, 42, 0, None))
    ha = host(a)
    assert np.array_equal(ha, host(b))
    assert ha.min() > 0.0 and ha.max() <= 1.0 and abs(ha.mean() - 0.5) < 2e-3 and abs(ha.var() - 1 / 12) < 2e-3
    # sharding independence: element i only depends on (seed, offset+i)
    ok(lib().t4k_rand_at(ptr(b), n // 2, t4.UNIFORM, 0.0, 1.0, 42, n // 2, None))
    assert np.array_equal(host(b)[: n // 2], ha[n // 2:])
    ok(lib().t4k_rand_at(ptr(a), n, t4.NORMAL, 0.0, 1.0, 7, 0, None))
    hn = host(a)
    assert abs(hn.mean()) < 5e-3 and abs(hn.std() - 1.0) < 5e-3
    # weight-init convention of Model::RAND: scale*2*(-0.5+U) in [-k, k)  (model.cpp:74-79)
    ok(lib().t4k_rand_at(ptr(a), n, t4.UNIFORM, -0.5, 2 * 0.1, 1, 0, None))
    hw = host(a)
    assert hw.min() > -0.1 - 1e-7 and hw.max() <= 0.1 + 1e-7


def test_launches_are_counted():
    before = lib().t4k_launch_count()
    d = zeros(1024)
    ok(lib().t4k_map(t4.FILL, ptr(d), 1.0, 1024, None))
    assert lib().t4k_launch_count() == before + 1


# ------------------------------------------------------------------ dataset feeding
@pytest.mark.parametrize("n", [0, 1, 15, 16, 784 * 5 + 3, 512 * 784])
@pytest.mark.parametrize("norm", [None, (128.0, 128.0), (0.0, 255.0)])
def test_dataset_load_bit_exact(n, norm):
    """Dataset::_load on device: U8 pixels -> (x - mean) * scale, labels U8 -> int32 -> one-hot; bit-exact vs the oracle"""
    rng = np.random.default_rng(n + 1)
    u8 = rng.integers(0, 256, max(n, 1), dtype=np.uint8)[:n]
    lab = rng.integers(0, 12, 37, dtype=np.uint8)
    mean, scale = orc.dataset_normalize(*norm) if norm else (np.float32(0.0), np.float32(1.0 / 256.0))
    src = torch.from_numpy(np.concatenate([u8, np.zeros(16, np.uint8)])).cuda()
    dst = zeros(max(n, 1)); l8 = torch.from_numpy(lab).cuda(); l32 = torch.zeros(37, dtype=torch.int32, device="cuda")
    hot2 = zeros(37, 10) + 7.0
    ok(lib().t4k_dataset_load(ptr(src), ptr(dst), n, float(mean), float(scale), C.c_void_p(l8.data_ptr()), C.c_void_p(l32.data_ptr()), 37, ptr(hot2), 10, None))
    assert_exact(host(hot2), orc.onehot(lab.astype(np.int32), 10), "one-hot written by the load launch")
    assert_exact(host(dst)[:n], orc.dataset_load(u8, mean, scale), "dataset_load")
    assert np.array_equal(l32.cpu().numpy(), lab.astype(np.int32))
    hot = zeros(37, 10)
    ok(lib().t4k_onehot(C.c_void_p(l32.data_ptr()), ptr(hot), 37, 10, None))
    assert_exact(host(hot), orc.onehot(lab.astype(np.int32), 10), "onehot of u8 labels (label >= E -> class 0, loss.cpp:66)")


@pytest.mark.parametrize("N,feedN,C0,hw", [(8, 8, 10, 28), (8, 5, 10, 28), (4, 4, 16, 28), (4, 3, 6, 28), (4, 4, 10, 18)])
def test_conv_pool_relu_fwd_feed_equals_load_then_block(N, feedN, C0, hw):
    """the forward block fed from a staged U8 mini-batch (t4k_conv_pool_relu_fwd_feed: Dataset::_load + Model::onehot + conv -> maxpool
    -> relu -> flatten in one launch) writes bit for bit what t4k_dataset_load followed by t4k_conv_pool_relu_fwd writes — dataset
    tensor (partial batch: the tail keeps its values), labels, one-hot rows, input copy and every layer tensor; C0 = 6 and 18 x 18
    (324 pixels: not a multiple of 16) take the entry point's unfused path"""
    rng = np.random.default_rng(N * 100 + feedN + C0)
    u8 = rng.integers(0, 256, (N, hw, hw, 1), dtype=np.uint8)
    lab = rng.integers(0, 12, N, dtype=np.uint8)                       # labels >= E fold to class 0 (loss.cpp:66)
    mean, scale = orc.dataset_normalize(128.0, 128.0)
    F, B = rnd(1, 3, 3, C0), rnd(C0)
    old = rnd(N, hw, hw, 1)                                            # what the dataset tensor held before (kept past feedN)
    E, hp = 10, hw // 2
    dF, dB = dev(F), dev(B)
    s8 = torch.from_numpy(np.concatenate([u8.ravel(), np.zeros(16, np.uint8)])).cuda(); l8 = torch.from_numpy(lab).cuda()

    def run(fused):
        data = dev(old); icopy = zeros(N, hw, hw, 1); l32 = torch.full((N,), -1, dtype=torch.int32, device="cuda"); hot = zeros(N, E) + 7.0
        cO, pO, aO, aF, fO = zeros(N, hw, hw, C0), zeros(N, hp, hp, C0), zeros(N, hp, hp, C0), zeros(N, hp, hp, C0), zeros(N, hp * hp * C0)
        u8p, l8p, l32p = C.c_void_p(s8.data_ptr()), C.c_void_p(l8.data_ptr()), C.c_void_p(l32.data_ptr())
        if fused:
            ok(lib().t4k_conv_pool_relu_fwd_feed(u8p, l8p, feedN, float(mean), float(scale), l32p, ptr(hot), E, ptr(data), ptr(dF), ptr(dB), ptr(icopy),
                                                 ptr(cO), ptr(pO), ptr(aO), ptr(aF), ptr(fO), N, hw, hw, 1, hw, hw, C0, 3, 1, 1, None), "fwd_feed")
        else:
            ok(lib().t4k_dataset_load(u8p, ptr(data), feedN * hw * hw, float(mean), float(scale), l8p, l32p, feedN, ptr(hot), E, None))
            ok(lib().t4k_conv_pool_relu_fwd(ptr(data), ptr(dF), ptr(dB), ptr(icopy), ptr(cO), ptr(pO), ptr(aO), ptr(aF), ptr(fO), N, hw, hw, 1, hw, hw, C0, 3, 1, 1, None))
        return [host(t).copy() for t in (data, icopy, hot, cO, pO, aO, aF, fO)] + [l32.cpu().numpy().copy()]
    a, b = run(True), run(False)
    for x, y, nm in zip(a, b, ("dataset tensor", "input copy", "one-hot", "conv", "pool", "relu", "mask", "flatten", "labels")):
        assert np.array_equal(x.view(np.uint32) if x.dtype == np.float32 else x, y.view(np.uint32) if y.dtype == np.float32 else y), nm
    want = np.concatenate([orc.dataset_load(u8[:feedN].ravel(), mean, scale), old.ravel()[feedN * hw * hw:]])
    assert_exact(a[0], want, "dataset tensor vs oracle (fed part) and previous values (tail)")
    assert_exact(a[2][:feedN], orc.onehot(lab[:feedN].astype(np.int32), E), "one-hot vs oracle")
    assert np.array_equal(a[8][:feedN], lab[:feedN].astype(np.int32)) and np.all(a[8][feedN:] == -1)


@pytest.mark.parametrize("layer", [t4.L_RELU, t4.L_TANH, t4.L_LEAKYRL])
@pytest.mark.parametrize("N,E1,EH,E0", [(512, 1960, 100, 10), (37, 300, 128, 32), (8, 64, 20, 3)])
def test_linear_act_head_fwd_equals_the_two_calls(layer, N, E1, EH, E0):
    """hidden linear + activation + classifier head with the split-K finish fused into the head kernel: bit-equal to
    t4k_linear_act_fwd followed by t4k_mlp_head_fwd (same sums, same order)"""
    X, W1, B1 = rnd(N, E1), rnd(EH, E1) * 0.05, rnd(EH)
    W2, B2 = rnd(E0, EH) * 0.3, rnd(E0)
    Xd, W1d, B1d, W2d, B2d = dev(X), dev(W1), dev(B1), dev(W2), dev(B2)
    y1a, a1a, f1a, y2a, pa = zeros(N, EH), zeros(N, EH), zeros(N, EH), zeros(N, E0), zeros(N, E0)
    y1b, a1b, f1b, y2b, pb, pd = zeros(N, EH), zeros(N, EH), zeros(N, EH), zeros(N, E0), zeros(N, E0), zeros(N, E0)
    ok(lib().t4k_linear_act_fwd(layer, ptr(Xd), ptr(W1d), ptr(B1d), ptr(y1a), ptr(a1a), ptr(f1a), 0.1, N, EH, E1, None))
    ok(lib().t4k_mlp_head_fwd(ptr(a1a), ptr(W2d), ptr(B2d), ptr(y2a), ptr(pa), N, E0, EH, None))
    ok(lib().t4k_linear_act_head_fwd(layer, ptr(Xd), ptr(W1d), ptr(B1d), ptr(y1b), ptr(a1b), ptr(f1b), 0.1,
                                     ptr(W2d), ptr(B2d), ptr(y2b), ptr(pb), ptr(pd), N, EH, E1, E0, None))
    for a, b, nm in ((y1a, y1b, "Y1"), (a1a, a1b, "A1"), (f1a, f1b, "F1"), (y2a, y2b, "Y2"), (pa, pb, "P"), (pa, pd, "Pdup")):
        assert_exact(host(b), host(a), nm)
    assert lib().t4k_linear_act_head_fwd(layer, ptr(Xd), ptr(W1d), ptr(B1d), ptr(y1b), ptr(a1b), ptr(f1b), 0.1,
                                         ptr(W2d), ptr(B2d), ptr(y2b), ptr(pb), None, N, 129, E1, E0, None) == t4.ENOSUP
