"""development aid: SIMT vs the tcgen05 engines vs the warp-MMA engine on the linear-layer GEMM shapes (MNIST CNN N=512, GAN N=1024)
and a few squares (engine crossover); graph-timed"""
import ctypes as C, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tensorforth_b200 import lib as t4
L = t4.load()
st = torch.cuda.Stream(); torch.cuda.set_stream(st); h = C.c_void_p(st.cuda_stream)
p = lambda t: C.c_void_p(t.data_ptr())


def gtime(fn, reps=20, replays=5):
    for _ in range(3): fn()
    st.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=st):
        for _ in range(reps): fn()
    g.replay(); st.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(st)
    for _ in range(replays): g.replay()
    b.record(st); st.synchronize()
    return a.elapsed_time(b) / (reps * replays) * 1e3


shapes = []   # (name, M, N, K, tA, tB)
for N, layers in ((512, [(1960, 100)]), (1024, [(784, 512), (512, 256), (128, 256), (256, 512), (512, 784)])):
    for E1, E0 in layers:
        shapes += [("N%d fwd  %d->%d" % (N, E1, E0), N, E0, E1, 0, 1), ("N%d dW   %d->%d" % (N, E1, E0), E0, E1, N, 1, 0), ("N%d dX   %d->%d" % (N, E1, E0), N, E1, E0, 0, 0)]
shapes += [("square 512", 512, 512, 512, 0, 0), ("square 1024", 1024, 1024, 1024, 0, 0), ("square 2048", 2048, 2048, 2048, 0, 0)]
engines = (("simt", t4.GEMM_SIMT), ("tc(pack)", t4.GEMM_TC), ("tcf", t4.GEMM_TCF), ("tl", t4.GEMM_TL), ("auto", t4.GEMM_AUTO))
for name, M, Nn, K, tA, tB in shapes:
    A = torch.randn((K, M) if tA else (M, K), device="cuda"); B = torch.randn((Nn, K) if tB else (K, Nn), device="cuda"); O = torch.zeros(M, Nn, device="cuda")
    ref = (A.double().T if tA else A.double()) @ (B.double().T if tB else B.double())
    gf = 2.0 * M * Nn * K / 1e9
    line = "GEMM %-22s M=%4d N=%4d K=%4d  %.2f GFLOP " % (name, M, Nn, K, gf)
    for en, eng in engines:
        try:
            rc = L.t4k_gemm_ex(eng, p(A), p(B), p(O), 1.0, 0.0, tA, tB, M, Nn, K, 1, 1, 0, 0, 0, h)
            us = gtime(lambda: L.t4k_gemm_ex(eng, p(A), p(B), p(O), 1.0, 0.0, tA, tB, M, Nn, K, 1, 1, 0, 0, 0, h)) if rc == 0 else float("nan")
            err = float(((O.double() - ref).pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()).cpu()) if rc == 0 else float("nan")
        except Exception:
            us, err = float("nan"), float("nan")
        line += "  %s %6.2f us (%5.1f TF/s, %.1e)" % (en, us, gf / us * 1e3, err)
    print(line, flush=True)
