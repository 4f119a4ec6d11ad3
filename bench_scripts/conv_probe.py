"""development aid for ncu: a few launches of the tcgen05 conv kernels at the config-5 shape (N=256 of 8192)"""
import ctypes as C, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tensorforth_b200 import lib as t4
L = t4.load(); p = lambda t: C.c_void_p(t.data_ptr())
cn = 256
f32 = lambda *s: torch.empty(*s, device="cuda").uniform_(-1, 1)
Ic, Fc, Bc, Oc, dXc, dFc, dBc = f32(cn, 56, 56, 64), f32(64, 3, 3, 64) * 0.1, f32(64), f32(cn, 56, 56, 64), f32(cn, 56, 56, 64), f32(64, 3, 3, 64), f32(64)
for _ in range(2):
    t4.check(L.t4k_conv2d_fwd(p(Ic), p(Fc), p(Bc), p(Oc), cn, 56, 56, 64, 56, 56, 64, 3, 1, 1, None))
    t4.check(L.t4k_conv2d_bwd(p(Ic), p(Oc), p(Fc), p(dXc), p(dFc), p(dBc), cn, 56, 56, 64, 56, 56, 64, 3, 1, 1, 1, None))
torch.cuda.synchronize(); print("ok")
