"""development aid (torchrun, one rank per GPU): per-call latency of the peer-store all-reduce (csrc/comm.cu) against NCCL,
each captured 20x into a CUDA graph and replayed, CUDA events on the launching stream, max over ranks."""
import ctypes as C, os, sys
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tensorforth_b200 import lib as t4, dp
rank, world, local = dp.env_rank()
torch.cuda.set_device(local)
L = t4.load()
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
sizes = [4, 4096, 197712, 1 << 20]
comm = dp.PeerComm(max(sizes))
h = C.c_void_p(st.cuda_stream)


def gtime(fn, reps=20, replays=10):
    for _ in range(3):
        fn()
    st.synchronize(); dist.barrier()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=st):
        for _ in range(reps):
            fn()
    g.replay(); st.synchronize(); dist.barrier(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(st)
    for _ in range(replays):
        g.replay()
    b.record(st); st.synchronize()
    t = torch.tensor([a.elapsed_time(b) / (reps * replays) * 1e3], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.cpu()[0])


for n in sizes:
    buf = torch.ones(n, device="cuda") * 1e-3
    ours = gtime(lambda: t4.check(L.t4k_allreduce_sum(comm.handle, C.c_void_p(buf.data_ptr()), n, h)))
    buf2 = torch.ones(n, device="cuda") * 1e-3
    nccl = gtime(lambda: dist.all_reduce(buf2))
    if rank == 0:
        print("EXCH world=%d variant=%s chunks=%s n=%d floats: peer-store %.2f us   nccl %.2f us" % (
            world, os.environ.get("T4K_COMM_VARIANT", "0"), os.environ.get("T4K_COMM_CHUNKS", "-"), n, ours, nccl), flush=True)
assert comm.status() == 0
dist.destroy_process_group()
