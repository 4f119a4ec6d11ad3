#!/usr/bin/env python
"""
bench.py — the measurement contract (task §④).

  python bench.py --gpus N --steps K --warmup W            our arm (N>1: launched by torch.distributed.run)
  python bench.py --impl reference --gpus N --steps K ...   the reference's own legacy CUDA build (oracle/_ref/ten4)

Metric (BASELINE.json): MNIST-CNN training samples/s — the CNN of examples/t4_40a.4th:10-13 at N=512 per GPU (weak scaling),
one step = forward + loss.ce + backprop + nn.adam on synthetic 28x28x1 data.  One JSON line on stdout (rank 0).

`value`   : device-timed (CUDA events on the launching stream, max over ranks), inputs resident in HBM, one CUDA-graph launch per step.
            The timed region is EXACTLY --steps steps between two events (barrier + synchronize on both sides); a 20-step window of this
            workload is 1.4 ms, so the window is repeated (up to 15 times) and the MEDIAN window is reported — every window is listed
            under `timing`.
`e2e`     : the same step through the public host API the way a training loop over a dataset runs: every step's mini-batch starts in
            pinned host memory as U8 pixels + U8 labels (what an MNIST loader holds), Dataset.stage() copies the bytes, the device
            normalises and one-hots them (inside the step's first kernel), and every step's loss is read back on the host (Model.train_step).
`roofline`: the call around the dominant kernel of the step (largest single launch of the committed ncu launch list, named in
            profiles/ncu_traffic.json), timed live (graph of 20 launches replayed between two events on its stream), against the
            measured HBM bandwidth; `calls` lists every call of the step the same way.
`extras`  : the other headline numbers — GEMM 4096^3 (both tensor-core engines, with their measured error), conv2d 3x3 64->64 @56x56 at
            the full N=8192 sharded over the ranks, one GAN iteration of examples/t4_40b.4th at N=1024 per GPU; N > 1: a strong-scaling
            line (the same global batch of 512 cut into N shards).
`cpu_baseline`: the C oracle (oracle/, "port") timed on the host cores on a bounded sample — reported, not a target.
N > 1     : one rank per GPU (torchrun); the gradient exchange is fused into the optimizer kernel over NVLink peer memory
            (`--exchange nccl` selects the NCCL all-reduce arm).
tensorForth has no CPU tensor path; the reference arm (`--impl reference`) therefore times the reference's own CUDA build on GPU 0
(SURVEY.md §8d).
"""
import argparse
import ctypes as C
import json
import os
import re
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

emit = None
BATCH = 512                       # per GPU (BASELINE config 3)
GAN_BATCH = 1024                  # per GPU (BASELINE config 4)
REF_TEN4 = os.path.join(ROOT, "oracle", "_ref", "ten4")


def peaks():
    p = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "src": "fallback"}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            m = json.load(f)
        p.update({k: m[k] for k in ("hbm_gbs", "bf16_tflops", "bf16_tflops_sustained") if k in m})
        p["src"] = "measured"
    except Exception:
        pass
    return p


# --------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.stop, self.t = index, [], threading.Event(), None

    def _run(self):
        while not self.stop.is_set():
            try:
                o = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                   capture_output=True, text=True, timeout=5).stdout.strip()
                if o:
                    self.rows.append([x.strip() for x in o.split(",")])
            except Exception:
                pass
            self.stop.wait(0.2)

    def __enter__(self):
        self.t = threading.Thread(target=self._run, daemon=True); self.t.start(); return self

    def __exit__(self, *a):
        self.stop.set(); self.t.join(timeout=3)

    def summary(self):
        sm = sorted(int(float(r[0])) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = max([int(float(r[1])) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()] or [0])
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(len(r) > 3 + k and r[3 + k].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": reasons, "samples": len(sm)}


WORKLOAD = ("MNIST CNN (examples/t4_40a.4th:10-13) conv3x3(1->10)+maxpool2+relu+flatten+linear100+relu+linear10+softmax, "
            "N=%d per GPU, step = forward + loss.ce + backprop + nn.adam(lr=1e-3)" % 512)


# --------------------------------------------------------------------------------------- reference arm
def ref_script(kind, warm, steps, batch):
    """Forth text both builds could run (SURVEY.md §8d 'synthetic step script')"""
    if kind == "mnist":
        return "\n".join([
            "0 trace", "%d constant N" % batch,
            "N 28 28 1 nn.model 0.5 10 conv2d 2 maxpool relu flatten 100 linear relu 10 linear softmax constant md0",
            "N 28 28 1 tensor rand 2 *= 1 -= constant X", "N 1 10 1 tensor rand constant Y",
            ": step ( M -- M ) X forward Y loss.ce drop Y backprop 0.001 nn.adam ;",
            ": bench ( M n -- M ) clock >r for step next clock r> - . ;",
            "md0 %d bench cr" % max(warm - 1, 0), "%d bench cr" % (steps - 1), "bye", ""])
    if kind == "gemm":
        return "\n".join([
            "0 trace", "4096 4096 matrix rand", "4096 4096 matrix rand",
            ": mx ( A B n -- A B ) clock >r for @ drop next clock r> - . ;",
            "%d mx cr" % max(warm - 1, 0), "%d mx cr" % (steps - 1), "bye", ""])
    if kind == "gan":                     # examples/t4_40b.4th:37-67 on a fixed synthetic "real" batch (no dataset files), no loss reads
        return "\n".join([
            "0 trace", "%d constant N" % batch,
            "N 1 1 1 tensor ones constant REAL", "N 1 1 1 tensor zeros constant FAKE",
            "N 28 28 1 nn.model 512 linear 0.2 leakyrelu 0.3 dropout 256 linear 0.2 leakyrelu 0.3 dropout 1 linear sigmoid constant D",
            "N 128 1 1 nn.model 256 linear 0.2 leakyrelu 512 linear 0.2 leakyrelu 784 linear tanh constant G",
            "N 28 28 1 tensor rand 2 *= 1 -= constant RX",
            ": X N 128 1 1 tensor randn ;",
            ": F G X forward -1 n@ N 28 28 1 reshape4 swap drop ;",
            ": train_d 1 trainable RX forward REAL backprop F forward FAKE backprop 0.0001 0.5 nn.adam ;",
            ": train_g 0 trainable F forward REAL backprop 0 n@ G swap backprop 0.0004 0.5 nn.adam drop ;",
            ": bench ( D n -- D ) clock >r for train_d train_g next clock r> - . ;",
            "D %d bench cr" % max(warm - 1, 0), "%d bench cr" % (steps - 1), "bye", ""])
    raise ValueError(kind)


def run_ref(kind, warm, steps, batch=BATCH, timeout=900):
    """run the reference's own CUDA build; returns ms for `steps` iterations (its own `clock` word, syncs included)"""
    if not os.path.exists(REF_TEN4):
        return None, "oracle/_ref/ten4 not built"
    env = dict(os.environ)
    try:
        p = subprocess.run([REF_TEN4], input=ref_script(kind, warm, steps, batch), capture_output=True, text=True, timeout=timeout, env=env, cwd=ROOT)
    except Exception as e:
        return None, "ten4 failed: %r" % (e,)
    nums = [float(x) for x in re.findall(r"^(-?\d+(?:\.\d+)?)\s*$", p.stdout, flags=re.M)]
    if len(nums) < 2:
        return None, "could not parse ten4 output: %s" % p.stdout[-300:].replace("\n", " | ")
    return nums[-1], None


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t0 = time.time()
    ms, why = run_ref("mnist", args.warmup, args.steps)
    line = {"impl": "reference", "metric": "mnist_cnn_train_samples_per_sec", "unit": "samples/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD,
                       "note": "reference = chochain/tensorForth's own CUDA kernels, unmodified, built for sm_100 (oracle/ref/build_ref.sh), "
                               "1 GPU (it has no multi-GPU and no CPU tensor path); timed with its own `clock` word incl. its per-kernel syncs"}}
    if ms is None:
        # no reference binary on this box: fall back to the CPU oracle port (bounded sample)
        v, cores, sample = cpu_port_baseline(max(2, min(args.steps, 10)))
        line.update({"value": v, "ms_per_step": BATCH / v * 1e3, "cpu_baseline": {"value": v, "unit": "samples/s", "cores": cores, "kind": "port", "sample": sample},
                     "e2e": {"value": v, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "reference_unavailable": why})
    else:
        v = BATCH * args.steps / (ms / 1e3) if ms > 0 else float("inf")
        line.update({"value": v, "ms_per_step": ms / args.steps,
                     "cpu_baseline": {"value": v, "unit": "samples/s", "cores": 1, "kind": "reference",
                                      "sample": "%d steps of N=%d on GPU 0 through oracle/_ref/ten4 (one host thread drives the GPU)" % (args.steps, BATCH)},
                     "e2e": {"value": v, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
        if not args.no_extras:
            gms, gwhy = run_ref("gemm", 3, 10)
            line["extras"] = {"gemm4096": ({"ms": gms / 10, "tflops": 2 * 4096 ** 3 / (gms / 10) / 1e9} if gms else {"unavailable": gwhy})}
            ams, awhy = run_ref("gan", 3, 10, batch=GAN_BATCH)
            line["extras"]["gan"] = ({"ms_per_iteration": ams / 10, "samples_per_s": GAN_BATCH * 10 / (ams / 1e3), "batch": GAN_BATCH,
                                      "note": "train_d + train_g of t4_40b.4th:60-67, fixed synthetic real batch"} if ams else {"unavailable": awhy})
    line["wall_s"] = round(time.time() - t0, 2)
    emit(line)


# --------------------------------------------------------------------------------------- CPU oracle baseline
def cpu_port_baseline(steps=4, batch=BATCH):
    """the oracle's Model restatement (C kernels, OpenMP) on the host cores: bounded sample of the same workload"""
    import numpy as np
    import ctypes
    from oracle import oracle as orc
    # torchrun exports OMP_NUM_THREADS=1 to its workers: the baseline is "the host cores of the box", so the OpenMP team is sized here
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        ctypes.CDLL("libgomp.so.1").omp_set_num_threads(int(cores))
    except OSError:
        cores = int(os.environ.get("OMP_NUM_THREADS", cores))
    om = orc.OracleModel(batch, 28, 28, 1, seed=1)
    om.add(orc.L_CONV, 10, 0.5, [3, 1, 1, 1]).add(orc.L_MAXPOOL, 2).add(orc.L_RELU).add(orc.L_FLATTEN)
    om.add(orc.L_LINEAR, 100, 1.0).add(orc.L_RELU).add(orc.L_LINEAR, 10, 1.0).add(orc.L_SOFTMAX)
    rng = np.random.default_rng(0)
    x = (rng.random((batch, 28, 28, 1), dtype=np.float32) * 2 - 1).astype(np.float32)
    y = orc.onehot(rng.integers(0, 10, batch), 10)

    def step():
        om.forward(x); om.loss(orc.LOSS_CE, y); om.backprop(y); om.adam(0.001)
    step()
    t0 = time.time()
    n = 0
    while n < steps or (time.time() - t0 < 10.0 and n < 200):
        step(); n += 1
    dt = time.time() - t0
    return batch * n / dt, cores, "%d train steps of N=%d (%.1f s) with the C oracle, OpenMP" % (n, batch, dt)


# --------------------------------------------------------------------------------------- conv sweep (BASELINE config 5)
def conv_extra(t4, L, torch, dist, rank, world, local, lib_stream, st, pk):
    """conv2d 3x3 s1 p1, NHWC 8192 x 56 x 56 x 64 -> 64, forward + backward, the 8192 samples sharded over the ranks (strong
    scaling: 8192 / world per GPU, no exchange in forward; backward sum-all-reduces dF, dB = 36 928 floats over the peer-store
    exchange).  3xTF32 implicit GEMM on tcgen05: tensor-bound; the HBM figure is the metric BASELINE.json names."""
    import ctypes as C
    from tensorforth_b200 import dp as t4dp
    NT = 8192
    cn = NT // world
    p = lambda t: C.c_void_p(t.data_ptr())
    f32 = lambda *s: torch.empty(*s, device="cuda").uniform_(-1, 1)
    Ic, Fc, Bc, Oc = f32(cn, 56, 56, 64), f32(64, 3, 3, 64) * 0.1, f32(64), f32(cn, 56, 56, 64)
    dXc = f32(cn, 56, 56, 64)
    dFB = torch.zeros(64 * 9 * 64 + 64, device="cuda")           # dF | dB contiguous: one exchange
    dFc, dBc = dFB[:64 * 9 * 64], dFB[64 * 9 * 64:]
    comm = t4dp.PeerComm(dFB.numel()) if world > 1 else None

    def kt(fn, iters):
        for _ in range(2):
            fn()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(lib_stream)
        for _ in range(iters):
            fn()
        b.record(lib_stream); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / iters
        if world > 1:
            t = torch.tensor([ms], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t.cpu()[0])
        return ms * 1e3

    def bwd():
        t4.check(L.t4k_conv2d_bwd(p(Ic), p(Oc), p(Fc), p(dXc), p(dFc), p(dBc), cn, 56, 56, 64, 56, 56, 64, 3, 1, 1, 1, st), "conv bwd")
        if comm is not None:
            t4.check(L.t4k_allreduce_sum(comm.handle, p(dFB), dFB.numel(), st), "dF exchange")
    usf = kt(lambda: t4.check(L.t4k_conv2d_fwd(p(Ic), p(Fc), p(Bc), p(Oc), cn, 56, 56, 64, 56, 56, 64, 3, 1, 1, st), "conv fwd"), 4)
    usb = kt(bwd, 3)
    if comm is not None:
        assert comm.status() == 0
    if rank != 0:
        return None
    n_el = NT * 56 * 56 * 64
    fb = (2 * n_el + Fc.numel()) * 4                           # whole job: read I once, write O once
    bb = (3 * n_el + 2 * Fc.numel()) * 4                       # read I, dO; write dX; dF
    fl = 2.0 * NT * 56 * 56 * 64 * 64 * 9
    tf32_peak = pk["bf16_tflops"] / 2
    return {"samples_total": NT, "samples_per_gpu": cn, "fwd_ms": round(usf / 1e3, 3), "bwd_ms": round(usb / 1e3, 3),
            "fwd_GBps": round(fb / usf / 1e3, 1), "bwd_GBps": round(bb / usb / 1e3, 1),
            "fwd_tflops": round(fl / usf / 1e6, 2), "bwd_tflops": round(2 * fl / usb / 1e6, 2),
            "roofline": {"bound": "hbm", "achieved": round(fb / usf / 1e3 / world, 1), "peak": pk["hbm_gbs"], "unit": "GB/s per GPU",
                         "frac": round(fb / usf / 1e3 / world / pk["hbm_gbs"], 4)},
            "roofline_tensor": {"bound": "tensor", "achieved": round(fl / usf / 1e6 / world, 1), "peak": round(tf32_peak / 3, 1), "unit": "TFLOP/s per GPU",
                                "frac": round(fl / usf / 1e6 / world / (tf32_peak / 3), 4)},
            "note": "whole-job numbers over %d GPU(s), max over ranks; 3xTF32 implicit GEMM (FP32-grade): tensor-bound, the HBM fraction is shown because "
                    "BASELINE.json names it; backward includes the dF/dB exchange" % world}


# --------------------------------------------------------------------------------------- GAN (BASELINE config 4)
def gan_extra(th, t4, L, torch, dist, rank, world, local, lib_stream, iters=50, warm=5):
    """train_d + train_g of examples/t4_40b.4th:60-67 at N=1024 per GPU on synthetic 28x28 data; data parallel: each model's
    gradient arena is exchanged inside its Adam kernel (csrc/comm.cu).  Eager launches (dropout draws a fresh mask every
    forward, so the iteration is not graph-captured).  Returns the extras entry (rank 0) or None."""
    import numpy as np
    from tensorforth_b200 import dp as t4dp
    N = GAN_BATCH
    L.t4k_rand_seed(4321)
    D, G = th.gan_discriminator(N, 0.3), th.gan_generator(N)
    rng = np.random.default_rng(200 + rank)
    real = th.Tensor.from_numpy((rng.random((N, 28, 28, 1), dtype=np.float32) * 2 - 1).astype(np.float32))
    REAL, FAKE = th.Tensor.tensor(N, 1, 1, 1, np.ones((N, 1), np.float32)), th.Tensor.tensor(N, 1, 1, 1, np.zeros((N, 1), np.float32))
    z1, z2 = th.Tensor.tensor(N, 128, 1, 1), th.Tensor.tensor(N, 128, 1, 1)
    exchange = "none"
    keep = []
    if world > 1:
        dev = torch.device("cuda", local)
        try:
            keep = [t4dp.DataParallel(D, dev, fused=True), t4dp.DataParallel(G, dev, fused=True)]
            exchange = "fused"
        except Exception as e:
            sys.stderr.write("rank %d: GAN fused exchange unavailable (%r)\n" % (rank, e))
            return {"unavailable": "peer exchange unavailable"} if rank == 0 else None

    def it():
        # the two `X` draws of an iteration; data parallel: every rank draws ITS shard of the global latent batch (per-rank Philox offset =
        # global element index, csrc/rand.cu) — and, through Model::dp_shard, its shard of the global dropout masks
        if world > 1:
            z1.rand_sharded(rank, world, normal=True); z2.rand_sharded(rank, world, normal=True)
        else:
            z1.randn(); z2.randn()
        th.gan_iteration(D, G, real, z1, z2, REAL, FAKE, losses=False)
    n0 = L.t4k_launch_count(); it(); launches = L.t4k_launch_count() - n0
    for _ in range(warm):
        it()

    def timed(fn):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(lib_stream)
        for _ in range(iters):
            fn()
        b.record(lib_stream)
        torch.cuda.synchronize()
        ms = a.elapsed_time(b)
        if world > 1:
            t = torch.tensor([ms], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t.cpu()[0])
        return ms
    ms_eager = timed(it)
    # the same iteration captured once (draws included: the graph starts with the RNG replay-epoch tick) and replayed
    ms = ms_eager
    graph_ok = True
    try:
        g = th.Graph(it)
        for _ in range(3):
            g()
        ms = timed(g)
    except Exception as e:
        graph_ok = False
        sys.stderr.write("rank %d: GAN graph capture unavailable (%r)\n" % (rank, e))
    if world > 1:
        for d in keep:
            assert d.comm.status() == 0
    l_dr, l_df, l_gr = th.gan_iteration(D, G, real, z1, z2, REAL, FAKE)
    if rank != 0:
        return None
    flops = 3 * 2 * N * (784 * 512 + 512 * 256 + 256) * 2 + 2 * N * (784 * 512 + 512 * 256 + 256) + 2 * 2 * N * (128 * 256 + 256 * 512 + 512 * 784) + 2 * 2 * N * (128 * 256 + 256 * 512 + 512 * 784)
    return {"batch_per_gpu": N, "ms_per_iteration": round(ms / iters, 4), "samples_per_s": round(N * world * iters / (ms / 1e3), 1),
            "launches_per_iteration": int(launches), "cuda_graph": graph_ok, "ms_per_iteration_eager": round(ms_eager / iters, 4),
            "exchange": exchange, "tflops": round(flops * world / (ms / iters) / 1e9, 2),
            "losses_after": {"d_real": round(l_dr, 4), "d_fake": round(l_df, 4), "g": round(l_gr, 4)},
            "note": "one iteration = train_d (D fwd/bwd on real + on G's fakes, Adam b1=0.5) + train_g (D frozen, dX of D's input through G, Adam); "
                    "3 D-forwards, 3 D-backwards, 2 G-forwards, 1 G-backward, 2 Adam; the whole iteration (latent draws and dropout masks included) "
                    "is one replayed CUDA graph; no loss reads in the timed loop"}


# --------------------------------------------------------------------------------------- our arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--eager", action="store_true", help="no CUDA-graph capture of the step")
    ap.add_argument("--exchange", default="fused", choices=["fused", "nccl"],
                    help="N>1: gradient exchange fused into the optimizer kernel over NVLink peer memory (default), or NCCL all-reduce")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    # the contract is ONE JSON line on stdout: anything a library prints (NCCL banner, torchrun notices) goes to stderr
    sys.stdout.flush()
    real_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    global emit
    emit = lambda line: (real_out.write(json.dumps(line) + "\n"), real_out.flush())
    if args.impl == "reference":
        return reference_arm(args)

    import numpy as np
    import torch
    from tensorforth_b200 import lib as t4
    from tensorforth_b200 import host as th

    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device: the product has no CPU path"
    torch.cuda.set_device(local)
    th.init(local)
    L, H = t4.load(), th.load()
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib_stream = torch.cuda.ExternalStream(th.stream(), device=local)       # events/NCCL ordered with the library's stream
    torch.cuda.set_stream(lib_stream)
    pk = peaks()

    # ---- model + synthetic data of BASELINE config 3 (weights: the model's own Philox init, per-rank identical seed)
    L.t4k_rand_seed(1234)
    m = th.mnist_cnn(BATCH)
    rng = np.random.default_rng(100 + rank)
    x8 = torch.from_numpy(rng.integers(0, 256, (BATCH, 28, 28, 1), dtype=np.uint8)).pin_memory()          # what an MNIST loader holds: U8 pixels
    xh = torch.from_numpy(((x8.numpy().astype(np.float32) - 128.0) * np.float32(1.0 / 128.0)).astype(np.float32)).pin_memory()   # `128 128 normalize` (t4_40b.4th:52)
    proj = np.random.default_rng(7).standard_normal((784, 10)).astype(np.float32)    # learnable synthetic labels: argmax of a fixed
    lab = (xh.numpy().reshape(BATCH, 784) @ proj).argmax(1)                           # random projection of the image (same rule on every rank)
    y8 = torch.from_numpy(lab.astype(np.uint8)).pin_memory()
    yh = torch.from_numpy(np.eye(10, dtype=np.float32)[lab]).pin_memory()
    X, Y = th.Tensor.tensor(BATCH, 28, 28, 1), th.Tensor.tensor(BATCH, 1, 10, 1)
    H.t4h_tensor_h2d(X.h, C.c_void_p(xh.data_ptr()), xh.numel()); H.t4h_tensor_h2d(Y.h, C.c_void_p(yh.data_ptr()), yh.numel())
    loss_dev = torch.zeros(8, device="cuda")
    lossp = C.c_void_p(loss_dev.data_ptr())
    LR = 1e-3

    from tensorforth_b200 import dp as t4dp
    dpm = None
    graph = not args.eager
    fused = False                 # data parallel: gradient exchange fused into the optimizer kernel (peer stores over NVLink)

    def step(Xt=None, Yt=None):
        Xt, Yt = Xt or X, Yt or Y
        if world == 1 or fused:
            # one CUDA-graph launch per step.  Data parallel: forward + loss + backprop on this rank's shard, then ONE
            # kernel pushes the flat gradient arena (and the loss sum) to every peer over NVLink, sums in rank order and runs Adam
            if graph:
                t4.check(m.step_graph(Xt, Yt, t4.LOSS_CE, lossp, optimizer=2, lr=LR), "step_graph")
            else:
                m.forward(Xt); m.loss_async(t4.LOSS_CE, Yt, lossp); m.backprop(Yt); m.adam(LR)
        else:
            # NCCL arm (--exchange nccl, or no peer access): graph(forward+loss+backprop) + NCCL SUM all-reduce + Adam launch
            if graph:
                t4.check(m.step_graph(Xt, Yt, t4.LOSS_CE, lossp, optimizer=-1, lr=LR), "step_graph")
            else:
                m.forward(Xt); m.loss_async(t4.LOSS_CE, Yt, lossp); m.backprop(Yt)
            dpm.allreduce_grads()
            m.adam(LR)

    # first step eagerly: builds the flat parameter arenas and sizes every workspace
    m.forward(X); m.loss_async(t4.LOSS_CE, Y, lossp); m.backprop(Y); m.adam(LR)
    n0 = L.t4k_launch_count()
    m.forward(X); m.loss_async(t4.LOSS_CE, Y, lossp); m.backprop(Y); m.adam(LR)
    launches_per_step = L.t4k_launch_count() - n0
    exchange = "none"
    if world > 1:
        exchange = "nccl"
        if args.exchange == "fused":
            try:
                dpm = t4dp.DataParallel(m, torch.device("cuda", local), fused=True, scalars=loss_dev[:1])   # broadcasts rank 0's parameters
                fused, exchange = True, "fused"
            except Exception as e:                                          # no cudaIpc / peer access on this box
                sys.stderr.write("rank %d: fused exchange unavailable (%r), using NCCL\n" % (rank, e))
            ok = torch.tensor([1.0 if fused else 0.0], device="cuda"); dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if fused and float(ok.cpu()[0]) == 0.0:                         # all ranks or none
                m.dp_attach(None); fused, exchange, dpm = False, "nccl", None
        if dpm is None:
            dpm = t4dp.DataParallel(m, torch.device("cuda", local))
    th.sync()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # kernels of ONE step as the timed loop runs it: the first graph step captures (every launch wrapper counts at capture time, a replay launches
    # exactly those nodes); the eager count above is what --eager would launch
    n0 = L.t4k_launch_count()
    step()
    if graph:
        launches_per_step = L.t4k_launch_count() - n0
    for _ in range(args.warmup):
        step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # The timed region is EXACTLY --steps steps between two events (barrier + synchronize on both sides, max over ranks).  A 20-step window
    # of this workload is 1.4 ms — one scheduling hiccup of the host thread is 5 % of it — so the window is repeated and the MEDIAN window is
    # the line's value; every window's time is in "timing".
    nwin = max(1, min(15, 300 // max(args.steps, 1)))

    def window(run):
        barrier()
        e0.record(lib_stream)
        run()
        e1.record(lib_stream)
        barrier()
        w = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([w], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); w = float(t.cpu()[0])
        return w

    def steps_k():
        for _ in range(args.steps):
            step()
    with ClockSampler(local) as cs:
        wins = sorted(window(steps_k) for _ in range(nwin))
        ms = wins[len(wins) // 2] if len(wins) & 1 else 0.5 * (wins[len(wins) // 2 - 1] + wins[len(wins) // 2])
        final_loss = float(loss_dev[0].cpu()) / (world if fused else 1)   # last timed step: global mean (fused: the loss sums ride in the exchange) or this rank's shard
        if ms < 600:                                           # keep the GPU under the same load so nvidia-smi sees clocks under load
            for _ in range(min(20000, int(800.0 / max(ms / args.steps, 1e-3)))):   # same count on every rank (collective inside)
                step()
            barrier()
    value = BATCH * world * args.steps / (ms / 1e3)

    # ---- e2e through the public host API, the way a training loop over a dataset runs (`ds for forward loss.ce backprop nn.adam next`):
    # EVERY step's mini-batch starts in pinned host memory as the U8 pixels + U8 labels a loader holds, and every step's loss is
    # read back on the host.  Dataset.stage() copies the bytes asynchronously (copy stream, double buffered: batch i+1 travels
    # while batch i trains); step_graph_ds() normalises them on the device (t4k_dataset_load), one-hots the labels and replays the
    # captured step; the loss of step i is read on the host while step i+1 runs.
    ds = th.Dataset(BATCH, 28, 28, 1).normalize(128.0, 128.0)
    lh = [torch.zeros(1).pin_memory() for _ in range(2)]
    lready = [torch.cuda.Event() for _ in range(2)]
    losses_seen = []

    def e2e_run(nsteps):
        ds.stage(x8, y8)
        if world == 1 or fused:
            # one host call per iteration: Model::train_step = commit (normalise + one-hot, 1 launch) + the captured step + async loss
            # D2H into pinned memory; it hands back the previous iteration's loss (read-back pipelined by one step)
            for i in range(nsteps):
                if i + 1 < nsteps:
                    ds.stage(x8, y8)
                prev = m.train_step_ds(ds, t4.LOSS_CE, lossp, optimizer=2, lr=LR)
                if i >= 1:
                    losses_seen.append(prev)
            losses_seen.append(m.train_flush())
            return
        for i in range(nsteps):                                   # NCCL arm
            b = i & 1
            if i + 1 < nsteps:
                ds.stage(x8, y8)
            t4.check(m.step_graph_ds(ds, t4.LOSS_CE, lossp, optimizer=-1, lr=LR), "step_graph_ds")
            dpm.allreduce_grads(); m.adam(LR)
            lh[b].copy_(loss_dev[:1], non_blocking=True); lready[b].record(lib_stream)
            if i >= 1:
                lready[b ^ 1].synchronize(); losses_seen.append(float(lh[b ^ 1][0]))  # step i-1's loss, on the host
        lready[(nsteps - 1) & 1].synchronize(); losses_seen.append(float(lh[(nsteps - 1) & 1][0]))

    e2e_run(max(args.warmup, 4))
    wins_e2e = []
    for _ in range(nwin):
        losses_seen.clear()
        wins_e2e.append(window(lambda: e2e_run(args.steps)))
        assert len(losses_seen) == args.steps
    e2e_last_loss = (losses_seen[-1] / (world if fused else 1)) if losses_seen else None      # what the host read for the last step of the last window
    wins_e2e.sort()
    ms_e2e = wins_e2e[len(wins_e2e) // 2] if len(wins_e2e) & 1 else 0.5 * (wins_e2e[len(wins_e2e) // 2 - 1] + wins_e2e[len(wins_e2e) // 2])
    e2e = {"value": BATCH * world * args.steps / (ms_e2e / 1e3), "unit": "samples/s",
           "h2d_bytes_per_step": int(x8.numel() + y8.numel()), "d2h_bytes_per_step": 4,
           "note": "per step: U8 pixels + U8 labels from pinned host memory -> async H2D (copy stream, double buffered) -> on-device normalise "
                   "(u8-128)/128 + one-hot, folded into the step's first kernel (t4k_conv_pool_relu_fwd_feed) -> train step -> loss D2H read on the host (pipelined by one step); one host call per iteration.  "
                   "It can come out a hair ABOVE `value`: the feed kernel reads 0.4 MB of U8 where the resident-input step copies 1.6 MB of FP32 into the model's input layer, "
                   "and the H2D copy + loss read-back overlap the step on their own streams",
           "window_ms": [round(w, 4) for w in wins_e2e], "last_loss_read_on_host": e2e_last_loss}

    # ---- strong scaling: the SAME global batch of 512 cut into world shards of 512/world samples, one data-parallel step
    strong = None
    if world > 1 and fused and BATCH % world == 0 and BATCH // world >= 32:
        try:
            nb = BATCH // world
            L.t4k_rand_seed(1234)
            ms_ = th.mnist_cnn(nb)
            Xs, Ys = th.Tensor.tensor(nb, 28, 28, 1), th.Tensor.tensor(nb, 1, 10, 1)
            H.t4h_tensor_h2d(Xs.h, C.c_void_p(xh.data_ptr()), nb * 784); H.t4h_tensor_h2d(Ys.h, C.c_void_p(yh.data_ptr()), nb * 10)
            ls_ = torch.zeros(8, device="cuda")
            ms_.forward(Xs); ms_.backprop(Ys); ms_.adam(LR); th.sync()
            dps = t4dp.DataParallel(ms_, torch.device("cuda", local), fused=True, scalars=ls_[:1])

            def sstep():
                t4.check(ms_.step_graph(Xs, Ys, t4.LOSS_CE, C.c_void_p(ls_.data_ptr()), optimizer=2, lr=LR), "step_graph (strong)")
            for _ in range(6):
                sstep()
            wins_s = sorted(window(lambda: [sstep() for _ in range(100)]) for _ in range(3))
            assert dps.comm.status() == 0
            strong = {"global_batch": BATCH, "batch_per_gpu": nb, "ms_per_step": round(wins_s[1] / 100, 6), "samples_per_s": round(BATCH * 100 / (wins_s[1] / 1e3), 1),
                      "scaling": "strong", "note": "the N=512 batch of the 1-GPU line cut into %d shards; same captured data-parallel step (exchange fused with Adam); "
                                                   "compare with the 1-GPU ms_per_step of the same box session" % world}
        except Exception as e:
            strong = {"unavailable": repr(e)[:200]}
            sys.stderr.write("strong-scaling extra failed: %r\n" % (e,))
    gan = conv = None
    if not args.no_extras:
        try:
            conv = conv_extra(t4, L, torch, dist, rank, world, local, lib_stream, C.c_void_p(th.stream()), pk)
        except Exception as e:
            conv = {"unavailable": repr(e)[:200]}
            sys.stderr.write("conv extra failed: %r\n" % (e,))
        torch.cuda.empty_cache()
        try:
            gan = gan_extra(th, t4, L, torch, dist, rank, world, local, lib_stream)
        except Exception as e:                                  # extras never take the headline line down
            gan = {"unavailable": repr(e)[:200]}
            sys.stderr.write("GAN extra failed: %r\n" % (e,))
    if fused:
        stt = dpm.comm.status()
        assert stt == 0, "rank %d: the gradient exchange timed out waiting for rank %d" % (rank, stt - 1)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- per-call live timing of the calls the step is made of, at the step's shapes → roofline of the dominant one.
    # CPU out of the loop: each call is captured 20x into a CUDA graph on a side stream and the graph replayed; CUDA events
    # on THAT stream bracket the replays (the step itself is a graph too, so this is the regime the kernels run in).
    def gtime(fn, reps=20, replays=10):
        side = torch.cuda.Stream(device=local)
        h = C.c_void_p(side.cuda_stream)
        with torch.cuda.stream(side):
            for _ in range(3):
                t4.check(fn(h), "probe")
            side.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=side):
                for _ in range(reps):
                    fn(h)
            g.replay(); side.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(side)
            for _ in range(replays):
                g.replay()
            b.record(side)
            side.synchronize()
        return a.elapsed_time(b) / (reps * replays) * 1e3        # us per call

    N = BATCH
    f32 = lambda *s: torch.empty(*s, device="cuda").uniform_(-1, 1)
    p = lambda t: C.c_void_p(t.data_ptr())
    I, F, Bv, I0, cO = f32(N, 28, 28, 1), f32(1, 3, 3, 10), f32(10), f32(N, 28, 28, 1), f32(N, 28, 28, 10)
    pO, aO, aF, fO, dY = (f32(N, 14, 14, 10) for _ in range(5))
    dXb, dF, dB = f32(N, 28, 28, 1), f32(1, 3, 3, 10), f32(10)
    W1, B1, Y1, A1, F1 = f32(100, 1960) * 0.05, f32(100), f32(N, 100), f32(N, 100), f32(N, 100)
    W2, B2, Y2, Pp, Tt = f32(10, 100), f32(10), f32(N, 10), f32(N, 10), f32(N, 10)
    dW1, dB1, dW2, dB2, dX1 = f32(100, 1960), f32(100), f32(10, 100), f32(10), f32(N, 1960)
    G_, DG_, M_, V_ = (f32(197710) for _ in range(4))
    lossd = torch.zeros(4, device="cuda")
    Pd = f32(N, 10)
    hscr = torch.zeros(max(int(L.t4k_head_train_scratch_floats(t4.L_RELU, N, 100, 1960, 10)), 4), device="cuda")
    hncta = C.c_int(0)
    fl = lambda *ts: sum(t.numel() for t in ts) * 4
    kernels = [  # (name, call, algorithmic bytes = every operand read once / every result written once)
        ("conv_pool_relu_fwd (+input copy, +flatten)", lambda h: L.t4k_conv_pool_relu_fwd(p(I), p(F), p(Bv), p(I0), p(cO), p(pO), p(aO), p(aF), p(fO), N, 28, 28, 1, 28, 28, 10, 3, 1, 1, h),
         fl(I, I0, cO, pO, aO, aF, fO)),
        ("linear_act_head_train 1960->100 relu ->10 softmax + the head's backward on the same rows (layer GEMM, tcgen05, mode 4)",
         lambda h: L.t4k_linear_act_head_train(t4.L_RELU, p(fO), p(W1), p(B1), p(Y1), p(A1), p(F1), 0.0, p(W2), p(B2), p(Y2), p(Pp), p(Pd), p(Tt), p(hscr), C.byref(hncta), N, 100, 1960, 10, h),
         fl(fO, W1, Y1, A1, F1, W2, Y2, Pp, Pd, Tt)),
        ("loss.ce (side stream)", lambda h: L.t4k_loss(t4.LOSS_CE, p(Pd), p(Tt), N * 10, N, p(lossd), h), fl(Pd, Tt)),
        ("head_grad_finish (dW2, dB2, dB1 from per-CTA partials; side stream)", lambda h: L.t4k_head_grad_finish(p(hscr), max(hncta.value, 1), 10, 100, p(dW2), p(dB2), p(dB1), h),
         hscr.numel() * 4 + fl(dW2, dW2, dB2, dB1)),
        ("linear_bwd_pair 1960->100 (dX1 and dW1 += in one launch of the layer GEMM)", lambda h: L.t4k_linear_bwd_pair(p(fO), p(W1), p(Y1), p(dX1), p(dW1), N, 100, 1960, h), fl(fO, Y1, dW1, dW1, Y1, W1, dX1)),
        ("conv_pool_relu_bwd (flatten', relu', pool', dF,dB,dX)", lambda h: L.t4k_conv_pool_relu_bwd(p(dY), p(aO), p(aF), p(pO), p(cO), p(I0), p(dXb), p(F), p(dF), p(dB), N, 28, 28, 1, 28, 28, 10, 3, 1, 1, 1, h),
         fl(dY, aF, cO, I0, aO, pO, cO, I0, dXb)),
        ("adam (197710 params)", lambda h: L.t4k_adam(p(G_), p(DG_), p(M_), p(V_), 1e-3, 0.9, 0.999, 197710, h), 7 * 197710 * 4),
    ]
    ktab = []
    for name, fn, nbytes in kernels:
        n0 = L.t4k_launch_count(); fn(None); nl = L.t4k_launch_count() - n0
        us = gtime(fn)
        ktab.append({"call": name, "launches": int(nl), "us": round(us, 2), "alg_MB": round(nbytes / 1e6, 2), "GBps": round(nbytes / us / 1e3, 1),
                     "hbm_frac": round(nbytes / us / 1e3 / pk["hbm_gbs"], 4)})
    # The dominant KERNEL of the step is the largest single launch of the committed ncu launch list (profiles/: k_cpr2_bwd, 18 % of the
    # step; each of the three GEMM launches is 11-12 %), named with its call in profiles/ncu_traffic.json; the calls of this table that
    # take several launches (linear_bwd: dW GEMM + split-K finish + dX GEMM, the first two on a side stream in the real step) are not
    # one kernel.  Without that file: the longest call.
    dom = max(ktab, key=lambda r: r["us"])
    traffic, traffic_src = None, None
    try:                                                       # dram__bytes_read+write per launch of the dominant kernel, from the committed ncu --set full capture
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            tj = json.load(f)
        named = [r for r in ktab if r["call"] == tj.get("call")]
        if named:
            dom = named[0]
            traffic, traffic_src = tj["dram_bytes_per_launch"], tj["source"]
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": dom["call"], "achieved": dom["GBps"], "peak": pk["hbm_gbs"], "unit": "GB/s",
                "frac": round(dom["GBps"] / pk["hbm_gbs"], 4), "traffic": traffic, "traffic_src": traffic_src, "peak_src": pk["src"] + " (burst copy bandwidth)",
                "alg_bytes_per_launch": int(dom["alg_MB"] * 1e6), "us_per_launch": dom["us"],
                "step_call_sum_us": round(sum(r["us"] for r in ktab), 1),
                "step_hbm_floor_us": round(sum(r["alg_MB"] for r in ktab) * 1e6 / (pk["hbm_gbs"] * 1e9) * 1e6, 1),
                "note": "dominant = the largest call of the step in this live, warm, graph-replayed timing (the conv block's backward: k_cpr2_bwd + its short finish "
                        "launch, timed together, so `achieved` is conservative for the kernel alone).  In the cold-cache serialised ncu list (profiles/r02_launches_step.txt) "
                        "the train-tail launch of the layer GEMM is the larger single kernel (24.7 % vs 20.3 %); it is latency-bound, not throughput-bound — its line is `layer_gemm`"}
    # the two layer-GEMM launches of the step against the tensor roofline (3xTF32: three TF32 MMAs per FP32 product): latency-bound by design
    tf32x3_peak = pk["bf16_tflops"] / 2 / 3
    lg = []
    for row, gflop in ((ktab[1], 2.0 * N * (1960 * 100 + 2 * 100 * 10) / 1e9), (ktab[4], 2.0 * 2.0 * N * 1960 * 100 / 1e9)):
        tf = gflop / row["us"] * 1e3                 # GFLOP per us = PFLOP/s; x 1e3 = TFLOP/s
        lg.append({"call": row["call"], "us": row["us"], "gflop": round(gflop, 4), "tflops": round(tf, 2), "peak": round(tf32x3_peak, 1),
                   "frac": round(tf / tf32x3_peak, 4), "hbm_frac": row["hbm_frac"]})
    roofline["layer_gemm"] = lg

    out = {"metric": "mnist_cnn_train_samples_per_sec", "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f32", "data": "synthetic",
           "config": {"workload": WORKLOAD,
                      "global_batch": BATCH * world, "parallelism": "dp%d" % world if world > 1 else "single",
                      "cuda_graph": bool(graph), "exchange": exchange,
                      "l2": "working set per step ~190 MB > 126 MB L2; no explicit flush (back-to-back steps is the workload)"},
           "clocks": cs.summary(), "e2e": e2e, "gpu_launches": int(launches_per_step * args.steps),
           "launches_per_step": int(launches_per_step), "final_loss": final_loss,
           "timing": {"windows": nwin, "steps_per_window": args.steps, "value_from": "median window", "window_ms": [round(w, 4) for w in wins]},
           "roofline": roofline, "calls": ktab}

    # ---- extras: the other two headline numbers of BASELINE.json (1 GPU only)
    if world == 1 and not args.no_extras:
        st = C.c_void_p(th.stream())

        def kt(fn, iters=50):                                  # large kernels: plain back-to-back launches on the library stream
            for _ in range(3):
                fn()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); a.record(lib_stream)
            for _ in range(iters):
                fn()
            b.record(lib_stream); torch.cuda.synchronize()
            return a.elapsed_time(b) / iters * 1e3             # us
        ex = {}
        n = 4096
        A, B_, Oo = f32(n, n), f32(n, n), f32(n, n)
        tf32_peak = pk["bf16_tflops"] / 2                      # TF32 dense = 1/2 BF16
        idx = torch.arange(0, n, 256, device="cuda")
        r64 = A[idx].double() @ B_.double()

        def gemm_line(engine, name, peak, peak_note):
            us = kt(lambda: L.t4k_gemm_ex(engine, p(A), p(B_), p(Oo), 1.0, 0.0, 0, 0, n, n, n, 1, 1, 0, 0, 0, st), iters=20)
            torch.cuda.synchronize()
            err = float((Oo[idx].double() - r64).pow(2).mean().sqrt() / r64.pow(2).mean().sqrt())
            tf = 2 * n ** 3 / us / 1e6
            return {"ms": round(us / 1e3, 4), "tflops": round(tf, 1), "engine": name, "rms_rel_err_vs_f64": float("%.2e" % err),
                    "roofline": {"bound": "tensor", "achieved": round(tf, 1), "peak": round(peak, 1), "unit": "TFLOP/s", "frac": round(tf / peak, 4), "note": peak_note}}
        # what `@` / t4k_gemm runs at this size: BF16x3 (FP32 operands split into bf16 hi + lo, three MMAs per product)
        ex["gemm4096"] = gemm_line(t4.GEMM_AUTO, "tcgen05 BF16x3 (pack + mma), selected by t4k_gemm for this size class", pk["bf16_tflops"] / 3,
                                   "peak = measured BF16 %.0f / 3 (three MMAs per FP32 product); the reference's FP32-FMA kernel has ~3.8e-6 rms "
                                   "accumulation error at K=4096, this engine 4e-6" % pk["bf16_tflops"])
        ex["gemm4096_3xtf32"] = gemm_line(t4.GEMM_TC, "tcgen05 3xTF32 (pack + mma)", tf32_peak / 3,
                                          "peak = measured BF16 %.0f /2 (TF32) /3 (three MMAs per product)" % pk["bf16_tflops"])
        del A, B_, Oo
        out["extras"] = ex
    if conv is not None:
        out.setdefault("extras", {})["conv2d_3x3_64"] = conv
    if gan is not None:
        out.setdefault("extras", {})["gan_t4_40b"] = gan
    if strong is not None:
        out.setdefault("extras", {})["mnist_strong_scaling"] = strong
    if not args.no_cpu_baseline:
        v, cores, sample = cpu_port_baseline()
        out["cpu_baseline"] = {"value": round(v, 1), "unit": "samples/s", "cores": cores, "kind": "port", "sample": sample}
    emit(out)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
