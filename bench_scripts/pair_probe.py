"""CTA-pair GEMM (k_gemm_tc2, T4K_GEMM_PAIR=1) against float64 rows and the single-CTA kernel: error + time"""
import ctypes as C, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tensorforth_b200 import lib as t4
L = t4.load()
p = lambda t: C.c_void_p(t.data_ptr())
def run(M, N, K, eng, iters=20):
    g = torch.Generator(device="cuda").manual_seed(1)
    A = torch.rand(M, K, device="cuda", generator=g) * 2 - 1; B = torch.rand(K, N, device="cuda", generator=g) * 2 - 1
    O = torch.full((M, N), float("nan"), device="cuda")
    rc = L.t4k_gemm_ex(eng, p(A), p(B), p(O), 1.0, 0.0, 0, 0, M, N, K, 1, 1, 0, 0, 0, None)
    torch.cuda.synchronize()
    idx = torch.arange(0, M, max(1, M // 16), device="cuda")
    ref = A[idx].double() @ B.double()
    err = float((O[idx].double() - ref).pow(2).mean().sqrt() / ref.pow(2).mean().sqrt())
    nanc = int(torch.isnan(O).sum())
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        L.t4k_gemm_ex(eng, p(A), p(B), p(O), 1.0, 0.0, 0, 0, M, N, K, 1, 1, 0, 0, 0, None)
    b.record(); torch.cuda.synchronize()
    us = a.elapsed_time(b) / iters * 1e3
    return rc, err, nanc, us, 2.0 * M * N * K / us / 1e6
for (M, N, K) in ((4096, 4096, 4096), (2048, 4096, 1024), (4096, 2048, 2048), (8192, 8192, 4096), (4096, 4000, 4096)):
    for eng, nm in ((t4.GEMM_TC_BF16X3, "bf16x3"), (t4.GEMM_TC, "3xtf32")):
        rc, err, nanc, us, tf = run(M, N, K, eng)
        print("pair=%s %5dx%5dx%5d %-7s rc=%d rms_err=%.2e nan=%d  %8.1f us  %6.1f TFLOP/s" % (os.environ.get("T4K_GEMM_PAIR", "0"), M, N, K, nm, rc, err, nanc, us, tf), flush=True)
