"""bring-up probe for the MN-major operands of the layer GEMM: with the other operand a one-hot selector, the output shows WHICH element
of the operand tile the tensor core fetched for every logical (m, k); tries a few descriptor encodings (t4k_set_gemm_tl knobs 3-7)"""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tensorforth_b200 import lib as t4
L = t4.load()
p = lambda t: C.c_void_p(t.data_ptr())
M = N = 128; K = 32
L.t4k_set_gemm_tl(2, 1)
SW128, SW128_A32 = 3, 4
configs = [("type1 atom32 LBO4096 SBO512 k1024", 1, SW128_A32, 4096, 512, 1024),
           ("type1 atom32 LBO512 SBO4096 k1024", 1, SW128_A32, 512, 4096, 1024),
           ("type1 atom32 LBO4096 SBO1024 k1024", 1, SW128_A32, 4096, 1024, 1024),
           ("type2 sw128  LBO4096 SBO1024 k1024", 2, SW128, 4096, 1024, 1024),
           ("type2 sw128  LBO1024 SBO4096 k1024", 2, SW128, 1024, 4096, 1024),
           ("type1 sw128  LBO4096 SBO512 k1024", 1, SW128, 4096, 512, 1024)]


def decode(name, tA, tB):
    if tA:      # A stored [K][M], value m + 128 k; B = selector [N][K]
        A = (torch.arange(M, device="cuda")[None, :] + 128 * torch.arange(K, device="cuda")[:, None]).float().contiguous()
        B = torch.zeros(N, K, device="cuda"); B[torch.arange(K), torch.arange(K)] = 1.0
        tB_ = 1
        exp = (torch.arange(M, device="cuda")[:, None] + 128 * torch.arange(K, device="cuda")[None, :]).float()      # O[m][n<32]
        O = torch.full((M, N), -1.0, device="cuda")
        rc = L.t4k_gemm_ex(t4.GEMM_TL, p(A), p(B), p(O), 1.0, 0.0, 1, tB_, M, N, K, 1, 1, 0, 0, 0, None)
        torch.cuda.synchronize()
        got = O[:, :K]
    else:       # B stored [K][N], value n + 128 k; A = selector [M][K]
        B = (torch.arange(N, device="cuda")[None, :] + 128 * torch.arange(K, device="cuda")[:, None]).float().contiguous()
        A = torch.zeros(M, K, device="cuda"); A[torch.arange(K), torch.arange(K)] = 1.0
        exp = (torch.arange(N, device="cuda")[None, :] + 128 * torch.arange(K, device="cuda")[:, None]).float()      # O[m<32][n]
        O = torch.full((M, N), -1.0, device="cuda")
        rc = L.t4k_gemm_ex(t4.GEMM_TL, p(A), p(B), p(O), 1.0, 0.0, 0, 0, M, N, K, 1, 1, 0, 0, 0, None)
        torch.cuda.synchronize()
        got = O[:K, :]
    okc = int((got == exp).sum()); tot = exp.numel()
    g = got.cpu().numpy(); e = exp.cpu().numpy()
    line = "%-40s %s rc=%d exact %4d/%d" % (name, "A(M-major)" if tA else "B(N-major)", rc, okc, tot)
    if okc != tot:
        bad = np.argwhere(g != e)[:6]
        def dec(v):
            v = int(round(float(v))); return "(x=%d,k=%d)" % (v % 128, v // 128) if 0 <= v < 4096 and abs(v - float(v)) < 1e-3 else "%.4g" % v
        line += "  e.g. " + " ".join("%s->%s" % (dec(e[i, j]), dec(g[i, j])) for i, j in bad)
        line += "  nonzero=%d" % int((g != 0).sum())
    print(line, flush=True)
    return okc == tot


for name, ty, swz, lbo, sbo, ks in configs:
    for k, v in ((3, ty), (4, swz), (5, lbo), (6, sbo), (7, ks)):
        L.t4k_set_gemm_tl(k, v)
    decode(name, 1, 1); decode(name, 0, 0)
