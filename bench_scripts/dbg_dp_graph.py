import ctypes as C, os, sys, time
os.environ["T4K_COMM_TIMEOUT_S"] = "5"
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from tensorforth_b200 import lib as t4, host as th
from oracle import oracle as orc
import test_gpu_dp_lanes as tl
world, N = 2, 16
mode = sys.argv[1] if len(sys.argv) > 1 else "thread"
rk = tl.Ranks(world, lambda: th.mnist_cnn(N), scal=True)
rng = np.random.default_rng(3)
for step in range(4):
    x = (rng.random((world * N, 28, 28, 1), dtype=np.float32) * 2 - 1).astype(np.float32); y = orc.onehot(rng.integers(0, 10, world * N), 10)
    xs, ys = tl.shards(x, world), tl.shards(y, world)
    ts = []
    rk.each(lambda r, m: ts.append((th.Tensor.from_numpy(xs[r]), th.Tensor.tensor(N, 1, 10, 1, ys[r]))))
    rk.each(lambda r, m: th.sync())
    def f(r, m):
        t0 = time.time()
        rc = m.step_graph(ts[r][0], ts[r][1], t4.LOSS_CE, C.c_void_p(rk.scal[r].data_ptr()), optimizer=2, lr=1e-3)
        print("  step", step, "rank", r, "issued rc", rc, "in %.3f s" % (time.time() - t0), flush=True)
    t0 = time.time()
    (rk.each_thread if mode == "thread" else rk.each)(f)
    rk.each(lambda r, m: th.sync())
    st = [rk.L.t4k_comm_status(h) for h in rk.comms]
    print("step", step, "done in %.2f s" % (time.time() - t0), "status", st, "loss", [float(s.cpu()[0]) for s in rk.scal], flush=True)
    if any(st):
        break
