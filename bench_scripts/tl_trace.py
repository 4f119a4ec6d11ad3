"""phase timeline of the layer GEMM (CTA 0 of each launch): clock64 stamps -> microseconds at the SM clock nvidia-smi reports"""
import ctypes as C, os, subprocess, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tensorforth_b200 import lib as t4
L = t4.load()
p = lambda t: C.c_void_p(t.data_ptr())
names = ["start", "setup done", "TMA0 issued", "tile0 landed", "lo0 done", "MMA0 issued", "TMA last issued", "tile last landed", "MMA last issued",
         "acc ready", "parked", "cluster barrier", "row0 reduced", "epilogue done", "exit"]
tr = torch.zeros(16, dtype=torch.int64, device="cuda")
shapes = [("MNIST fwd", 512, 100, 1960, 0, 1), ("MNIST dW", 100, 1960, 512, 1, 0), ("MNIST dX", 512, 1960, 100, 0, 0),
          ("GAN fwd 784->512", 1024, 512, 784, 0, 1), ("GAN dW 784->512", 512, 784, 1024, 1, 0), ("GAN dX 512->256", 1024, 512, 256, 0, 0)]
import itertools
for (name, M, N, K, tA, tB), l2 in itertools.product(shapes, (1, 0)):
    L.t4k_set_gemm_tl(8, l2)
    A = torch.randn((K, M) if tA else (M, K), device="cuda"); B = torch.randn((N, K) if tB else (K, N), device="cuda"); O = torch.zeros(M, N, device="cuda")
    for _ in range(5):
        L.t4k_gemm_ex(t4.GEMM_TL, p(A), p(B), p(O), 1.0, 0.0, tA, tB, M, N, K, 1, 1, 0, 0, 0, None)
    torch.cuda.synchronize()
    L.t4k_gemm_tl_trace(p(tr))
    for _ in range(3):
        L.t4k_gemm_ex(t4.GEMM_TL, p(A), p(B), p(O), 1.0, 0.0, tA, tB, M, N, K, 1, 1, 0, 0, 0, None)
    torch.cuda.synchronize()
    L.t4k_gemm_tl_trace(None)
    mhz = float(subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm", "--format=csv,noheader,nounits"], capture_output=True, text=True).stdout.split()[0])
    t = tr.cpu().tolist()
    print("%s  M=%d N=%d K=%d  split-K partials through %s (SM clock %.0f MHz)" % (name, M, N, K, "L2" if l2 else "DSMEM", mhz))
    print("   " + "  ".join("%s %.2f" % (names[i], (t[i] - t[0]) / mhz) for i in range(15) if t[i]))
