import ctypes as C, os, sys, time
os.environ["T4K_COMM_TIMEOUT_S"] = "6"
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from tensorforth_b200 import lib as t4, host as th
from oracle import oracle as orc
import test_gpu_dp_lanes as tl
world, N = 2, 16
rk = tl.Ranks(world, lambda: th.mnist_cnn(N), scal=True)
L, H = rk.L, th.load()
rng = np.random.default_rng(3)
x = (rng.random((world * N, 28, 28, 1), dtype=np.float32) * 2 - 1).astype(np.float32); y = orc.onehot(rng.integers(0, 10, world * N), 10)
xs, ys = tl.shards(x, world), tl.shards(y, world)
ts = []
rk.each(lambda r, m: ts.append((th.Tensor.from_numpy(xs[r]), th.Tensor.tensor(N, 1, 10, 1, ys[r]))))
rk.each(lambda r, m: th.sync())
a, b = torch.ones(64, device="cuda"), torch.zeros(64, device="cuda")
torch.cuda.synchronize()
# rank 0: whole step (its exchange kernel will spin for rank 1)
th.use_lane(0)
m0 = rk.models[0]
m0.forward(ts[0][0]); m0.backprop(ts[0][1]); m0.adam(1e-3)
time.sleep(0.2)
th.use_lane(1)
s_main, s_side = th.stream(), H.t4h_side_stream()
def probe(name, fn, st):
    t0 = time.time(); fn()
    es = torch.cuda.ExternalStream(st)
    while not es.query() and time.time() - t0 < 1.0:
        time.sleep(0.01)
    print("%-50s %s after %.2f s" % (name, "DONE" if es.query() else "BLOCKED", time.time() - t0), flush=True)
    return es.query()
m1 = rk.models[1]
probe("lane1 main: copy", lambda: L.t4k_copy(C.c_void_p(a.data_ptr()), C.c_void_p(b.data_ptr()), 64, C.c_void_p(s_main)), s_main)
probe("lane1 side: copy", lambda: L.t4k_copy(C.c_void_p(a.data_ptr()), C.c_void_p(b.data_ptr()), 64, C.c_void_p(s_side)), s_side)
okf = probe("lane1 main: model forward", lambda: m1.forward(ts[1][0]), s_main)
if okf:
    okb = probe("lane1 main: model backprop", lambda: m1.backprop(ts[1][1]), s_main)
    probe("lane1 side after backprop", lambda: None, s_side)
    if okb:
        probe("lane1 main: adam (exchange)", lambda: m1.adam(1e-3), s_main)
torch.cuda.synchronize()
print("status", [L.t4k_comm_status(h) for h in rk.comms])
