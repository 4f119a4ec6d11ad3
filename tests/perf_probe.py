"""quick device-timed probes of individual C-ABI calls (development aid; bench.py is the contract)"""
import ctypes as C
import sys
import os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tensorforth_b200 import lib as t4

L = t4.load()


def p(t):
    return C.c_void_p(t.data_ptr())


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def gtime(fn, reps=20, replays=10):
    """device time per call with the CPU out of the loop: `reps` calls captured into one CUDA graph, replayed.
    fn(stream_handle) must issue its work on that stream."""
    st = torch.cuda.Stream()
    h = C.c_void_p(st.cuda_stream)
    with torch.cuda.stream(st):
        for _ in range(3):
            fn(h)
        st.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st):
            for _ in range(reps):
                fn(h)
        g.replay(); st.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(replays):
            g.replay()
        e1.record(st)
        st.synchronize()
    return e0.elapsed_time(e1) / (reps * replays)


def gemm(n, engine=t4.GEMM_AUTO):
    A = torch.rand(n, n, device="cuda") - 0.5
    B = torch.rand(n, n, device="cuda") - 0.5
    O = torch.zeros(n, n, device="cuda")
    ms = timeit(lambda: t4.check(L.t4k_gemm_ex(engine, p(A), p(B), p(O), 1.0, 0.0, 0, 0, n, n, n, 1, 1, 0, 0, 0, None)))
    print("gemm %d^3 engine=%d: %.3f ms  %.1f TFLOP/s" % (n, engine, ms, 2 * n ** 3 / ms / 1e9))
    ms = timeit(lambda: torch.matmul(A, B, out=O))
    print("   torch fp32 matmul (cuBLAS, reference point only): %.3f ms  %.1f TFLOP/s" % (ms, 2 * n ** 3 / ms / 1e9))


def stream(n=1 << 28):
    a = torch.rand(n, device="cuda"); b = torch.empty_like(a)
    ms = timeit(lambda: t4.check(L.t4k_copy(p(a), p(b), n, None)))
    print("copy %d MiB: %.3f ms  %.0f GB/s" % (n * 4 >> 20, ms, 2 * n * 4 / ms / 1e6))
    ms = timeit(lambda: t4.check(L.t4k_map(t4.RELU, p(a), 0.0, n, None)))
    print("map relu in place: %.3f ms  %.0f GB/s" % (ms, 2 * n * 4 / ms / 1e6))
    o = torch.zeros(4, device="cuda")
    ms = timeit(lambda: t4.check(L.t4k_sum(p(a), n, p(o), None)))
    print("sum: %.3f ms  %.0f GB/s" % (ms, n * 4 / ms / 1e6))
    ms = timeit(lambda: b.copy_(a))
    print("   torch copy_: %.3f ms  %.0f GB/s" % (ms, 2 * n * 4 / ms / 1e6))


def conv(n=256, c=64, hw=56, engine=t4.GEMM_AUTO):
    I = torch.rand(n, hw, hw, c, device="cuda") - 0.5
    F = (torch.rand(c, 3, 3, c, device="cuda") - 0.5) * 0.2
    B = torch.rand(c, device="cuda")
    O = torch.empty(n, hw, hw, c, device="cuda"); dX = torch.empty_like(I)
    dF = torch.zeros_like(F); dB = torch.zeros_like(B)
    L.t4k_set_conv_engine(engine)
    fl = 2.0 * n * hw * hw * c * c * 9
    by = 2 * I.numel() * 4
    ms = timeit(lambda: t4.check(L.t4k_conv2d_fwd(p(I), p(F), p(B), p(O), n, hw, hw, c, hw, hw, c, 3, 1, 1, None)), iters=5)
    print("conv fwd N=%d engine=%d: %.3f ms  %.1f TFLOP/s  %.0f GB/s" % (n, engine, ms, fl / ms / 1e9, by / ms / 1e6))
    ms = timeit(lambda: t4.check(L.t4k_conv2d_bwd(p(I), p(O), p(F), p(dX), None, None, n, hw, hw, c, hw, hw, c, 3, 1, 1, 0, None)), iters=5)
    print("conv dX  N=%d engine=%d: %.3f ms  %.1f TFLOP/s  %.0f GB/s" % (n, engine, ms, fl / ms / 1e9, by / ms / 1e6))
    ms = timeit(lambda: t4.check(L.t4k_conv2d_bwd(p(I), p(O), p(F), p(dX), p(dF), p(dB), n, hw, hw, c, hw, hw, c, 3, 1, 1, 1, None)), iters=5)
    print("conv bwd (dF,dB,dX) N=%d engine=%d: %.3f ms  %.1f TFLOP/s" % (n, engine, ms, 2 * fl / ms / 1e9))
    L.t4k_set_conv_engine(t4.GEMM_AUTO)


def cpr(n=512, c0=10, hw=28):
    f32 = lambda *s: torch.empty(*s, device="cuda").uniform_(-1, 1)
    I, F, B = f32(n, hw, hw, 1), f32(1, 3, 3, c0), f32(c0)
    I0 = torch.empty_like(I)
    cO, pO, aO, aF, fO = f32(n, hw, hw, c0), f32(n, hw // 2, hw // 2, c0), f32(n, hw // 2, hw // 2, c0), f32(n, hw // 2, hw // 2, c0), f32(n, hw // 2, hw // 2, c0)
    dY, dXb, dF, dB = f32(n, hw // 2, hw // 2, c0), torch.empty_like(I), torch.zeros_like(F), torch.zeros_like(B)
    fwd = lambda: t4.check(L.t4k_conv_pool_relu_fwd(p(I), p(F), p(B), p(I0), p(cO), p(pO), p(aO), p(aF), p(fO), n, hw, hw, 1, hw, hw, c0, 3, 1, 1, None))
    ms = timeit(fwd, iters=50)
    by = (2 * I.numel() + cO.numel() + 4 * pO.numel()) * 4
    print("cpr fwd N=%d: %.2f us  %.0f GB/s (alg %.1f MB)" % (n, ms * 1e3, by / ms / 1e6, by / 1e6))
    def bwd():
        t4.check(L.t4k_conv_pool_relu_bwd(p(dY), p(aO), p(aF), p(pO), p(cO), p(I0), p(dXb), p(F), p(dF), p(dB), n, hw, hw, 1, hw, hw, c0, 3, 1, 1, 1, None))
    fwd(); ms = timeit(bwd, iters=50)
    by = (dY.numel() * 4 + 2 * cO.numel() + 3 * I.numel()) * 4
    print("cpr bwd N=%d: %.2f us  %.0f GB/s (alg %.1f MB)" % (n, ms * 1e3, by / ms / 1e6, by / 1e6))


def fc(n=512):
    """the FC-tail kernels of the MNIST step at the step's shapes, back-to-back (warm L2) timing"""
    f32 = lambda *s: torch.empty(*s, device="cuda").uniform_(-1, 1)
    X, W1, B1 = f32(n, 1960), f32(100, 1960) * 0.05, f32(100)
    Y1, A1, F1 = f32(n, 100), f32(n, 100), f32(n, 100)
    W2, B2, Y2, P, T = f32(10, 100), f32(10), f32(n, 10), f32(n, 10), f32(n, 10)
    dW1, dB1, dW2, dB2, dX = f32(100, 1960), f32(100), f32(10, 100), f32(10), f32(n, 1960)
    loss = torch.zeros(4, device="cuda")
    tab = [
        ("linear_act_fwd 1960->100 (+bias+relu)", lambda h: L.t4k_linear_act_fwd(t4.L_RELU, p(X), p(W1), p(B1), p(Y1), p(A1), p(F1), 0.0, n, 100, 1960, h)),
        ("mlp_head_fwd 100->10 (+softmax)", lambda h: L.t4k_mlp_head_fwd(p(A1), p(W2), p(B2), p(Y2), p(P), n, 10, 100, h)),
        ("loss.ce", lambda h: L.t4k_loss(t4.LOSS_CE, p(P), p(T), n * 10, n, p(loss), h)),
        ("mlp_head_bwd", lambda h: L.t4k_mlp_head_bwd(p(P), p(T), p(Y2), p(A1), p(F1), p(Y1), p(W2), p(dW2), p(dB2), p(dB1), n, 10, 100, 1, h)),
        ("linear_bwd 1960->100 (dW,dX)", lambda h: L.t4k_linear_bwd_ex(p(X), p(W1), p(Y1), p(dX), p(dW1), p(dB1), n, 100, 1960, 1, 1, h)),
        ("gemm dW only", lambda h: L.t4k_gemm(p(Y1), p(X), p(dW1), 1.0, 1.0, 1, 0, 100, 1960, n, 1, 1, 0, 0, 0, h)),
        ("gemm dX only", lambda h: L.t4k_gemm(p(Y1), p(W1), p(dX), 1.0, 0.0, 0, 0, n, 1960, 100, 1, 1, 0, 0, 0, h)),
        ("gemm fwd only (no bias)", lambda h: L.t4k_gemm(p(X), p(W1), p(Y1), 1.0, 0.0, 0, 1, n, 100, 1960, 1, 1, 0, 0, 0, h)),
    ]
    I, F, B = f32(n, 28, 28, 1), f32(1, 3, 3, 10), f32(10)
    I0, cO = torch.empty_like(I), f32(n, 28, 28, 10)
    pO, aO, aF, fO, dY = (f32(n, 14, 14, 10) for _ in range(5))
    dXb, dF, dB = torch.empty_like(I), torch.zeros_like(F), torch.zeros_like(B)
    G, DG, Mm, V = (f32(197710) for _ in range(4))
    tab += [
        ("conv_pool_relu_fwd", lambda h: L.t4k_conv_pool_relu_fwd(p(I), p(F), p(B), p(I0), p(cO), p(pO), p(aO), p(aF), p(fO), n, 28, 28, 1, 28, 28, 10, 3, 1, 1, h)),
        ("conv_pool_relu_bwd (+wgrad fin)", lambda h: L.t4k_conv_pool_relu_bwd(p(dY), p(aO), p(aF), p(pO), p(cO), p(I0), p(dXb), p(F), p(dF), p(dB), n, 28, 28, 1, 28, 28, 10, 3, 1, 1, 1, h)),
        ("adam (one tensor, 197710 params)", lambda h: L.t4k_adam(p(G), p(DG), p(Mm), p(V), 1e-3, 0.9, 0.999, 197710, h)),
    ]
    tot = 0.0
    for name, fn in tab:
        n0 = L.t4k_launch_count(); t4.check(fn(None)); k = L.t4k_launch_count() - n0
        ms = gtime(lambda h: t4.check(fn(h)))
        print("%-42s %7.2f us  (%d launches)" % (name, ms * 1e3, k))


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what in ("all", "gemm"):
        for n in (1024, 2048, 4096, 8192):
            gemm(n, t4.GEMM_TC)
        gemm(1024, t4.GEMM_SIMT); gemm(4096, t4.GEMM_SIMT)
    if what in ("all", "conv"):
        conv(256)
        conv(64, engine=t4.GEMM_SIMT)
    if what in ("all", "fc"):
        fc()
    if what in ("all", "cpr"):
        cpr()
    if what in ("all", "stream"):
        stream()
