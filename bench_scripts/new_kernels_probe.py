"""development aid for ncu: one launch each of the kernels added late in round 1 — the single-launch tensor-core GEMM (GAN D layer 1
forward shape), the dataset load (MNIST batch of 512) and the fused exchange+Adam kernel on a 1-rank communicator (no peers:
its local traffic and instruction mix; the NVLink part needs N>1 and is timed by bench_scripts/exchange_probe.py)."""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tensorforth_b200 import lib as t4
L = t4.load()
p = lambda t: C.c_void_p(t.data_ptr())
M, N, K = 1024, 512, 784
A, B, O = torch.randn(M, K, device="cuda"), torch.randn(N, K, device="cuda"), torch.zeros(M, N, device="cuda")
for _ in range(3):
    t4.check(L.t4k_gemm_ex(t4.GEMM_TCF, p(A), p(B), p(O), 1.0, 0.0, 0, 1, M, N, K, 1, 1, 0, 0, 0, None))
u8 = torch.randint(0, 256, (512 * 784,), dtype=torch.uint8, device="cuda"); l8 = torch.randint(0, 10, (512,), dtype=torch.uint8, device="cuda")
dst, l32, hot = torch.zeros(512 * 784, device="cuda"), torch.zeros(512, dtype=torch.int32, device="cuda"), torch.zeros(512, 10, device="cuda")
for _ in range(3):
    t4.check(L.t4k_dataset_load(p(u8), p(dst), 512 * 784, 128.0, 1 / 128.0, p(l8), p(l32), 512, p(hot), 10, None))
h = C.c_void_p(); t4.check(L.t4k_comm_create(0, 1, 197712, C.byref(h), None)); t4.check(L.t4k_comm_connect(h, None))
seg = torch.from_numpy(np.array([[0, 197712, 1]], dtype=np.int64)).cuda()
G, DG, Mm, V = (torch.randn(197712, device="cuda") for _ in range(4))
for _ in range(3):
    t4.check(L.t4k_optim_multi_dp(h, 1, p(G), p(DG), p(Mm), p(V.abs_()), p(seg), 1, 197712, 1e-3, 0.9, 0.999, 0.0, None, 0, 0, None))
torch.cuda.synchronize()
print("ok")
