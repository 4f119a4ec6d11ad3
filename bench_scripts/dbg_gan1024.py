"""why does dw4 of the generator at N=1024 leave the 1e-4 band at step 1 (sgd)?  error of the layer GEMM relative to the float64 product and to
sum |a||b| (the scale FP32 rounding acts on), next to the oracle's own distance from float64"""
import ctypes as C, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from tensorforth_b200 import lib as t4, host as th
from oracle import oracle as orc
import test_gpu_model as tm
N = 1024
rng = np.random.default_rng(11)
gm, om, shape, E, lop = tm.build_pair("gan_g", N)
for step in range(2):
    x = (rng.random(shape, dtype=np.float32) * 2 - 1).astype(np.float32)
    y = (rng.random((N, E), dtype=np.float32) * 2 - 1).astype(np.float32)
    X, Y = th.Tensor.from_numpy(x), th.Tensor.tensor(N, 1, E, 1, y)
    gm.forward(X); om.forward(x)
    x4g, x4o = gm.layer(4).numpy().reshape(N, -1).copy(), om.layers[4].data.reshape(N, -1).copy()
    gm.backprop(Y); om.backprop(y)
    dyg, dyo = gm.layer(5).numpy().reshape(N, -1).copy(), om.layers[5].data.reshape(N, -1).copy()
    got, ref = gm.dw(4).numpy().reshape(784, 512), om.layers[4].dw.reshape(784, 512)
    f64 = dyo.astype(np.float64).T @ x4o.astype(np.float64)
    f64g = dyg.astype(np.float64).T @ x4g.astype(np.float64)
    sabs = np.abs(dyo).astype(np.float64).T @ np.abs(x4o).astype(np.float64)
    print("step", step, "x4 rms", np.sqrt((x4o**2).mean()), "max", np.abs(x4o).max(), "dy rms", np.sqrt((dyo**2).mean()), "max", np.abs(dyo).max())
    print("  operand diff gpu vs oracle: x4", np.abs(x4g - x4o).max(), "dy", np.abs(dyg - dyo).max())
    print("  dw rms", np.sqrt((ref**2).mean()), "max|got-ref|", np.abs(got - ref).max(), "max|ref-f64|", np.abs(ref - f64).max(), "max|got-f64(own operands)|", np.abs(got - f64g).max(),
          "max|f64g-f64|", np.abs(f64g - f64).max())
    print("  relative to sum|a||b|: got", (np.abs(got - f64g) / sabs).max(), "oracle", (np.abs(ref - f64) / sabs).max(), "sabs max", sabs.max())
    gm.sgd(0.05, 0.9); om.sgd(0.05, 0.9)
    for i, L in enumerate(om.layers[:-1]):
        if L.dw is not None and L.w is not None:
            L.w[...] = gm.w(i).numpy().reshape(L.w.shape); L.b[...] = gm.b(i).numpy().reshape(L.b.shape)
