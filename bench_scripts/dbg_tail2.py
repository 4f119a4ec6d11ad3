import ctypes as C, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from tensorforth_b200 import lib as t4, host as th
from oracle import oracle as orc
import test_gpu_model as tm
N = 32
rng = np.random.default_rng(4)
x = (rng.random((N, 28, 28, 1), dtype=np.float32) * 2 - 1).astype(np.float32); y = orc.onehot(rng.integers(0, 10, N), 10)
ga, om, *_ = tm.build_pair("mnist", N)
gb, _, *_ = tm.build_pair("mnist", N)
X, Y = th.Tensor.from_numpy(x), th.Tensor.tensor(N, 1, 10, 1, y)
la, lb = torch.zeros(1, device="cuda"), torch.zeros(1, device="cuda")
ga.forward(X); ga.loss_async(t4.LOSS_CE, Y, C.c_void_p(la.data_ptr())); th.sync()
pa = ga.layer(-1).numpy().copy()
ga.backprop(Y); th.sync()
rc = gb.step_graph(X, Y, t4.LOSS_CE, C.c_void_p(lb.data_ptr()), optimizer=-1, lr=0.001); th.sync()
print("rc", rc, "loss eager", float(la.cpu()[0]), "graph", float(lb.cpu()[0]))
for i in range(len(ga)):
    a, b = ga.layer(i).numpy(), gb.layer(i).numpy()
    print("layer", i, a.shape, "max diff", float(np.abs(a - b).max()), "rms a", float(np.sqrt((a * a).mean())))
for i in (0, 4, 6):
    print("dw", i, float(np.abs(ga.dw(i).numpy() - gb.dw(i).numpy()).max()), "db", float(np.abs(ga.db(i).numpy() - gb.db(i).numpy()).max()), "rms", float(np.sqrt((ga.dw(i).numpy() ** 2).mean())))
# --- the C-ABI train tail on the model's own tensors
L = t4.load()
gc_, _, *_ = tm.build_pair("mnist", N)
gc_.forward(X); th.sync()
inp = gc_.layer(4).numpy().reshape(N, 1960).copy(); pfw = gc_.layer(8).numpy().reshape(N, 10).copy()
W1, B1, W2, B2 = gc_.w(4).numpy().reshape(100, 1960).copy(), gc_.b(4).numpy().copy(), gc_.w(6).numpy().reshape(10, 100).copy(), gc_.b(6).numpy().copy()
dev = lambda a: torch.from_numpy(np.ascontiguousarray(a, np.float32)).cuda()
p = lambda t: C.c_void_p(t.data_ptr())
z = lambda *s: torch.zeros(*s, device="cuda")
Xd, W1d, B1d, W2d, B2d, Td = dev(inp), dev(W1), dev(B1), dev(W2), dev(B2), dev(y)
y1b, a1b, f1b, y2b, pb, pdb = z(N, 100), z(N, 100), z(N, 100), z(N, 10), z(N, 10), z(N, 10)
nf = L.t4k_head_train_scratch_floats(t4.L_RELU, N, 100, 1960, 10)
scratch = z(int(nf)); ncta = C.c_int(0)
rc2 = L.t4k_linear_act_head_train(t4.L_RELU, p(Xd), p(W1d), p(B1d), p(y1b), p(a1b), p(f1b), 0.0, p(W2d), p(B2d), p(y2b), p(pb), p(pdb), p(Td), p(scratch), C.byref(ncta), N, 100, 1960, 10, None)
torch.cuda.synchronize()
print("C-ABI on model tensors: rc", rc2, "Pdup vs model forward P: max diff", float(np.abs(pdb.cpu().numpy() - pfw).max()), "B1 shape", B1.shape, "bias sample", B1[:3], "in rms", float(np.sqrt((inp**2).mean())))
