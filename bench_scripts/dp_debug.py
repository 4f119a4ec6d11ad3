"""development aid: loss trajectory of the data-parallel MNIST step, eager vs CUDA-graph, under torchrun"""
import ctypes as C, os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tensorforth_b200 import lib as t4, host as th, dp
rank, world, local = dp.env_rank()
torch.cuda.set_device(local); th.init(local)
L, H = t4.load(), th.load()
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ls = torch.cuda.ExternalStream(th.stream(), device=local); torch.cuda.set_stream(ls)
for mode in ("eager", "graph"):
    L.t4k_rand_seed(1234)
    m = th.mnist_cnn(512)
    rng = np.random.default_rng(100 + rank)
    X = th.Tensor.from_numpy((rng.random((512, 28, 28, 1), dtype=np.float32) * 2 - 1))
    Y = th.Tensor.tensor(512, 1, 10, 1, np.eye(10, dtype=np.float32)[rng.integers(0, 10, 512)])
    loss_dev = torch.zeros(8, device="cuda"); lp = C.c_void_p(loss_dev.data_ptr())
    m.forward(X); m.loss_async(t4.LOSS_CE, Y, lp); m.backprop(Y); m.adam(1e-3)
    d = dp.DataParallel(m, torch.device("cuda", local)) if world > 1 else None
    out = []
    for i in range(60):
        if mode == "graph":
            t4.check(m.step_graph(X, Y, t4.LOSS_CE, lp, optimizer=-1 if d else 2, lr=1e-3), "g")
        else:
            m.forward(X); m.loss_async(t4.LOSS_CE, Y, lp); m.backprop(Y)
        if d:
            d.allreduce_grads(); m.adam(1e-3)
        elif mode == "eager":
            m.adam(1e-3)
        if i % 10 == 9:
            out.append(round(float(loss_dev[0].cpu()), 4))
    gn = float(d.grads.abs().sum().cpu()) if d else -1
    print("rank", rank, mode, out, "DG abs sum after step", gn, "err:", H.t4h_last_error(), flush=True)
if world > 1:
    dist.destroy_process_group()
