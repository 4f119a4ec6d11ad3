"""first contact of the layer GEMM (gemm_tl.cu) with a device: every transposition, with the raw plane as hi and with the masked hi stored
explicitly, a few cluster sizes; prints the error against float64 (run under `timeout`: a wrong barrier protocol hangs)"""
import ctypes as C, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tensorforth_b200 import lib as t4
L = t4.load()
p = lambda t: C.c_void_p(t.data_ptr())
torch.manual_seed(0)
bad = 0
for (M, N, K) in ((128, 128, 32), (128, 128, 256), (512, 100, 1960), (100, 1960, 512), (512, 1960, 100), (1024, 512, 784)):
    for tA in (0, 1):
        for tB in (0, 1):
            if ((M if tA else K) % 4) or ((K if tB else N) % 4):
                continue
            A = torch.rand((K, M) if tA else (M, K), device="cuda") * 2 - 1
            B = torch.rand((N, K) if tB else (K, N), device="cuda") * 2 - 1
            ref = (A.double().T if tA else A.double()) @ (B.double().T if tB else B.double())
            for mask in (0, 1):
                for smax in (16, 1):
                    L.t4k_set_gemm_tl(1, mask); L.t4k_set_gemm_tl(2, smax)
                    O = torch.full((M, N), float("nan"), device="cuda")
                    rc = L.t4k_gemm_ex(t4.GEMM_TL, p(A), p(B), p(O), 1.0, 0.0, tA, tB, M, N, K, 1, 1, 0, 0, 0, None)
                    torch.cuda.synchronize()
                    err = float(((O.double() - ref).pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()).cpu()) if rc == 0 else float("nan")
                    flag = "" if (rc == 0 and err < 1e-5) else "   <-- BAD"
                    bad += 1 if flag else 0
                    print("M=%4d N=%4d K=%4d tA=%d tB=%d mask_hi=%d smax=%2d rc=%d rms_err=%.2e%s" % (M, N, K, tA, tB, mask, smax, rc, err, flag), flush=True)
L.t4k_set_gemm_tl(1, 0); L.t4k_set_gemm_tl(2, 16)
print("bad:", bad)
sys.exit(1 if bad else 0)
