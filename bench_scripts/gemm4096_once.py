"""GEMM 4096^3 on both tensor-core engines, two calls each, for `ncu --set full -k regex:k_gemm_tc|k_pack`"""
import ctypes as C, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tensorforth_b200 import lib as t4
L = t4.load()
p = lambda t: C.c_void_p(t.data_ptr())
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
A, B, O = (torch.rand(n, n, device="cuda") * 2 - 1 for _ in range(3))
for eng in (t4.GEMM_TC_BF16X3, t4.GEMM_TC):
    for _ in range(2):
        t4.check(L.t4k_gemm_ex(eng, p(A), p(B), p(O), 1.0, 0.0, 0, 0, n, n, n, 1, 1, 0, 0, 0, None), "gemm")
    torch.cuda.synchronize()
print("done")
