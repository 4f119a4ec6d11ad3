"""
make_golden.py — generates tests/golden/ref_kernels.npz by running the REFERENCE's own CUDA kernels
(oracle/_ref/refkern = oracle/ref/refkern.cu linked with /root/reference/src/t4math.cu and
src/nn/nmath.cu, built by oracle/ref/build_ref.sh) on seeded inputs, on a B200:

    gpurun -- 'python tests/golden/make_golden.py gpurun_out/ref_kernels.npz'
    cp gpurun_out/ref_kernels.npz tests/golden/

The fixture holds inputs AND reference outputs, so neither the GPU box at test time nor the
CPU tests need /root/reference.  It pins (a) the CPU oracle (tests/test_oracle_golden.py, CPU)
and (b) the CUDA path (tests/test_gpu_golden.py, -m gpu) for the ops the reference's own
example scripts give no numbers for: conv2d, pool, softmax, batchnorm, Adam, ...
"""
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import refkern as rk          # noqa: E402
from oracle import oracle as orc          # noqa: E402  (enum values only)

rng = np.random.default_rng(20261017)


def rnd(*shape, lo=-1.0, hi=1.0):
    return (rng.random(shape, dtype=np.float32) * (hi - lo) + lo).astype(np.float32)


def main(out_path):
    G = {}
    recs = []
    keys = []

    def add(key, op, ints=(), flts=(), arrs=(), outs=()):
        """queue one reference kernel call; inputs stored as key/in{k}, outputs as key/{name}"""
        for k, a in enumerate(arrs):
            G["%s/in%d" % (key, k)] = np.asarray(a, np.float32)
        G["%s/ints" % key] = np.asarray(ints, np.int64)
        G["%s/flts" % key] = np.asarray(flts, np.float64)
        recs.append((op, ints, flts, arrs))
        keys.append((key, outs))

    # ---- GEMM (k_gemm_tile_claude via FORK3T) and the other variants
    for tA in (0, 1):
        for tB in (0, 1):
            M, N, K = 70, 50, 90
            A = rnd(K, M) if tA else rnd(M, K)
            B = rnd(N, K) if tB else rnd(K, N)
            add("gemm3_t%d%d" % (tA, tB), "gemm", [3, tA, tB, M, N, K, 1], [0.5, 2.0], [A, B, rnd(M, N)], ["O"])
    add("gemm3_c2", "gemm", [3, 0, 0, 33, 20, 17, 2], [1.0, 0.0], [rnd(33, 17, 2), rnd(17, 20, 2), np.zeros((33, 20, 2))], ["O"])
    add("gemm3_256", "gemm", [3, 0, 0, 256, 256, 256, 1], [1.0, 0.0], [rnd(256, 256), rnd(256, 256), np.zeros((256, 256))], ["O"])
    for v in (1, 2, 4):
        add("gemm%d" % v, "gemm", [v, 0, 0, 40, 30, 50, 1], [1.0, 0.5], [rnd(40, 50), rnd(50, 30), rnd(40, 30)], ["O"])
    # ---- elementwise
    for op, v in [(orc.ABS, 0), (orc.NEG, 0), (orc.EXP, 0), (orc.LN, 0), (orc.LOG, 0), (orc.TANH, 0), (orc.RELU, 0),
                  (orc.SIGM, 0), (orc.SQRT, 0), (orc.RCP, 0), (orc.SAT, 0), (orc.FILL, 3.25), (orc.GFILL, 2.0),
                  (orc.SCALE, 1.5), (orc.POW, 2.5), (orc.ADD, .75), (orc.SUB, .75), (orc.MUL, -3.0), (orc.DIV, 7.0)]:
        a = rnd(1000, lo=-2, hi=2)
        if op in (orc.POW, orc.RCP):
            a = np.abs(a) + 0.1
        add("map_%d" % op, "map", [op], [v], [a], ["A"])
    for op in (orc.ADD, orc.SUB, orc.MUL, orc.DIV):
        add("ts_%d" % op, "ts_op", [op], [1.7], [rnd(777), np.zeros(777)], ["O"])
        add("tt_%d" % op, "tt_op", [op], [], [rnd(777), rnd(777, lo=.5, hi=2), np.zeros(777)], ["O"])
    add("transpose", "transpose", [7, 5, 3], [], [rnd(7, 5, 3), np.zeros((5, 7, 3))], ["T"])
    # ---- reductions
    x = rnd(100000)
    add("sum", "sum", [], [], [x, np.zeros(1)], ["v"])
    add("nvar", "nvar", [], [0.25], [x, np.zeros(1)], ["v"])
    add("max", "max", [1], [], [x, np.full(1, -3.4e38)], ["v"])
    add("min", "max", [0], [], [x, np.full(1, 3.4e38)], ["v"])
    add("dot", "dot", [1000, 3], [0.5, 2.0], [rnd(1000, 3), rnd(1000, 3), rnd(3)], ["O"])
    add("bce", "bce", [], [], [(rng.random(5000) < .5).astype(np.float32), rnd(5000, lo=.01, hi=.99), np.zeros(1)], ["v"])
    # ---- nn small kernels
    add("bias", "bias", [64, 10], [], [rnd(10), rnd(64, 10)], ["Y"])
    add("dlinear_db", "dlinear_db", [64, 10], [], [rnd(64, 10), rnd(10)], ["dB"])
    for layer, alpha in [(orc.L_RELU, 0), (orc.L_TANH, 0), (orc.L_SIGMOID, 0), (orc.L_SELU, 0), (orc.L_LEAKYRL, .2),
                         (orc.L_ELU, 1.0), (orc.L_DROPOUT, .3)]:
        add("act_%d" % layer, "activate", [layer], [alpha], [rnd(2000, lo=-3, hi=3), np.zeros(2000), rng.random(2000, dtype=np.float32)], ["O", "F"])
    add("softmax_10", "softmax", [64, 10], [], [rnd(64, 10, lo=-4, hi=4), np.zeros((64, 10))], ["O"])
    add("softmax_300", "softmax", [8, 300], [], [rnd(8, 300, lo=-4, hi=4), np.zeros((8, 300))], ["O"])
    # ---- conv2d fwd/bwd, all four (K,S,P) configs + the MNIST first layer
    for name, (N, H, W, C1, C0, K, S, P) in {
        "k1": (2, 9, 9, 2, 4, 1, 1, 0), "k3": (2, 12, 12, 3, 5, 3, 1, 1), "k4": (2, 14, 14, 2, 6, 4, 2, 1),
        "k5": (2, 12, 12, 3, 5, 5, 1, 2), "mnist": (4, 28, 28, 1, 10, 3, 1, 1), "c16": (2, 14, 14, 16, 32, 3, 1, 1)}.items():
        H0 = W0 = (H - K + 2 * P) // S + 1
        I, F, B = rnd(N, H, W, C1), rnd(C1, K, K, C0, lo=-.3, hi=.3), rnd(C0)
        dims = [N, H, W, C1, H0, W0, C0, K, S, P]
        add("conv_" + name, "conv2d", dims, [], [I, F, B, np.zeros((N, H0, W0, C0))], ["O"])
        add("dconv_" + name, "dconv2d", dims + [1], [], [I, rnd(N, H0, W0, C0), F, np.zeros_like(I), rnd(C1, K, K, C0), rnd(C0)], ["dX", "dF", "dB"])
    # ---- pooling
    for layer in (orc.L_AVGPOOL, orc.L_MAXPOOL, orc.L_MINPOOL):
        for K, (N, H, W, Cc) in ((2, (3, 8, 8, 10)), (3, (2, 9, 12, 5))):
            x = rnd(N, H, W, Cc); x[0, :K, :K, 0] = 0.5
            add("pool_%d_%d" % (layer, K), "pool", [layer, N, H, W, H // K, W // K, Cc, K], [], [x, np.zeros((N, H // K, W // K, Cc))], ["O"])
            add("dpool_%d_%d" % (layer, K), "dpool", [layer, N, H, W, H // K, W // K, Cc, K], [], [x, rnd(N, H // K, W // K, Cc)], ["I"])
    add("upsample_2", "dpool", [orc.L_USAMPLE, 2, 8, 8, 4, 4, 3, 2], [], [np.zeros((2, 8, 8, 3)), rnd(2, 4, 4, 3)], ["I"])
    # ---- batchnorm
    N, H, W, Cc = 4, 5, 7, 6
    x, g, b = rnd(N, H, W, Cc, lo=-2, hi=3), rnd(Cc, lo=.5, hi=1.5), rnd(Cc)
    add("bn", "batchnorm", [N, H, W, Cc], [], [x, g, b, np.zeros_like(x), np.zeros_like(x), np.zeros(3 * Cc)], ["O", "XH", "scr"])
    # ---- optimizers
    n = 3000
    add("sgd0", "sgd", [3], [0.5, 0.0], [rnd(n), rnd(n), rnd(n)], ["G", "DG", "M"])
    add("sgdm", "sgd", [1], [0.5, 0.9], [rnd(n), rnd(n), rnd(n, lo=-.1, hi=.1)], ["G", "DG", "M"])
    add("adam", "adam", [1], [1e-3, .9, .999], [rnd(n), rnd(n), rnd(n, lo=-.1, hi=.1), rnd(n, lo=0, hi=.1)], ["G", "DG", "M", "V"])
    add("adamw", "adamw", [1], [1e-3, .9, .999, .01], [rnd(n), rnd(n), rnd(n, lo=-.1, hi=.1), rnd(n, lo=0, hi=.1)], ["G", "DG", "M", "V"])

    outs = rk.run(recs)
    for (key, names), arrs in zip(keys, outs):
        assert len(names) == len(arrs), key
        for nm, a in zip(names, arrs):
            G["%s/%s" % (key, nm)] = a
    # dbatchnorm needs the forward's xhat / rvar: second pass
    xh, scr = G["bn/XH"].reshape(x.shape), G["bn/scr"]
    dy, dW0, dB0 = rnd(N, H, W, Cc), rnd(Cc), rnd(Cc)
    r = rk.one("dbatchnorm", [N, H, W, Cc, 1], [], [dy, xh, g, dW0, dB0, scr, np.zeros_like(x)])
    for k, a in enumerate([dy, xh, g, dW0, dB0, scr]):
        G["dbn/in%d" % k] = np.asarray(a, np.float32)
    G["dbn/ints"] = np.asarray([N, H, W, Cc, 1], np.int64)
    G["dbn/dX"], G["dbn/dW"], G["dbn/dB"] = r
    np.savez_compressed(out_path, **G)
    print("wrote %s: %d arrays, %.1f KiB" % (out_path, len(G), os.path.getsize(out_path) / 1024))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "ref_kernels.npz"))
