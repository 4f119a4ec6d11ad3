"""
tests/golden/ref_kernels.npz holds inputs + outputs of the REFERENCE's own CUDA kernels run on a
B200 (generator: tests/golden/make_golden.py through oracle/_ref/refkern).  Two uses:

  * CPU (`-m "not gpu"`): the oracle must reproduce every reference output → pins the oracle for
    conv2d / pool / softmax / batchnorm / Adam ..., which the reference's own scripts give no
    numbers for.
  * GPU (`-m gpu`): the CUDA path, called through the C-ABI, must reproduce them too.

Tolerances: bit-exact for routing/index/IEEE elementwise ops; 1e-4 relative (north star) where the
reference's own result depends on atomicAdd order or on fast-math intrinsics.
"""
import ctypes as C
import os
import numpy as np
import pytest

from oracle import oracle as orc

HERE = os.path.dirname(os.path.abspath(__file__))
G = np.load(os.path.join(HERE, "golden", "ref_kernels.npz"))
KEYS = sorted({k.split("/")[0] for k in G.files})


def gin(key, k):
    return G["%s/in%d" % (key, k)]


def ints(key):
    return [int(x) for x in G[key + "/ints"]]


def flts(key):
    return [float(x) for x in G[key + "/flts"]]


EXACT_MAP = {orc.ABS, orc.NEG, orc.RELU, orc.SAT, orc.FILL, orc.SCALE, orc.ADD, orc.SUB, orc.MUL, orc.DIV, orc.SQRT, orc.RCP}


def close(got, ref, rtol=1e-4, atol=None, exact=False, what=""):
    got = np.asarray(got, np.float32).ravel(); ref = np.asarray(ref, np.float32).ravel()
    assert got.size == ref.size, (what, got.size, ref.size)
    if exact:
        assert np.array_equal(got.view(np.uint32), ref.view(np.uint32)), what
        return
    g, r = got.astype(np.float64), ref.astype(np.float64)
    if atol is None:
        atol = rtol * (np.sqrt(np.mean(r * r)) + 1e-30)
    bad = np.abs(g - r) > rtol * np.abs(r) + atol
    assert not bad.any(), "%s: %d/%d off, max abs err %.3e" % (what, bad.sum(), r.size, np.abs(g - r).max())


# ---------------------------------------------------------------------------------------
# one evaluator per op family, parameterised by a backend: ORC (CPU oracle) or GPU (C-ABI)
# each returns {output name: array} and a dict of exact-flags
# ---------------------------------------------------------------------------------------
class Orc:
    name = "oracle"

    def gemm(self, key):
        v, tA, tB, M, N, K, Cc = ints(key); al, be = flts(key)
        A, B, O = gin(key, 0), gin(key, 1), gin(key, 2)
        if v in (1, 2):
            o = orc.f32(O).copy()
            orc.lib().orc_gemm_f64acc(orc._p(orc.f32(A)), orc._p(orc.f32(B)), orc._p(o), al, be, M, N, K, Cc)
            return {"O": o}
        return {"O": orc.gemm(A, B, O, al, be, bool(tA), bool(tB), M, N, K, Cc)}

    def map(self, key):      return {"A": orc.map_(ints(key)[0], gin(key, 0), flts(key)[0])}
    def ts_op(self, key):    return {"O": orc.ts_op(ints(key)[0], gin(key, 0), flts(key)[0])}
    def tt_op(self, key):    return {"O": orc.tt_op(ints(key)[0], gin(key, 0), gin(key, 1))}
    def transpose(self, key): return {"T": orc.transpose(gin(key, 0), ints(key)[2])}
    def sum(self, key):      return {"v": [orc.tsum(gin(key, 0))]}
    def nvar(self, key):     return {"v": [float(orc.lib().orc_nvar(orc._p(orc.f32(gin(key, 0))), flts(key)[0], gin(key, 0).size))]}
    def minmax(self, key):   return {"v": [orc.tmax(gin(key, 0)) if ints(key)[0] else orc.tmin(gin(key, 0))]}
    def dot(self, key):      return {"O": orc.dot(gin(key, 0), gin(key, 1), gin(key, 2), *flts(key), C_=ints(key)[1])}
    def bce(self, key):      return {"v": [float(orc.lib().orc_bce_sum(orc._p(orc.f32(gin(key, 0))), orc._p(orc.f32(gin(key, 1))), gin(key, 0).size))]}

    def bias(self, key):
        y = orc.f32(gin(key, 1)).copy(); orc.lib().orc_bias(orc._p(orc.f32(gin(key, 0))), orc._p(y), *ints(key)); return {"Y": y}

    def dlinear_db(self, key):
        d = orc.f32(gin(key, 1)).copy(); orc.lib().orc_dlinear_db(orc._p(orc.f32(gin(key, 0))), orc._p(d), *ints(key)); return {"dB": d}

    def activate(self, key):
        o, f = orc.activate(ints(key)[0], gin(key, 0), flts(key)[0], mask=gin(key, 2)); return {"O": o, "F": f}

    def softmax(self, key):  return {"O": orc.softmax(gin(key, 0), ints(key)[0])}

    def conv2d(self, key):
        N, H1, W1, C1, H0, W0, C0, K, S, P = ints(key)
        return {"O": orc.conv2d(gin(key, 0), gin(key, 1), gin(key, 2), K, S, P)}

    def dconv2d(self, key):
        N, H1, W1, C1, H0, W0, C0, K, S, P, tr = ints(key)
        dX, dF, dB = orc.dconv2d(gin(key, 0), gin(key, 1), gin(key, 2), K, S, P, gin(key, 4), gin(key, 5), bool(tr))
        return {"dX": dX, "dF": dF, "dB": dB}

    def pool(self, key):     return {"O": orc.pool(ints(key)[0], gin(key, 0), ints(key)[7])}
    def dpool(self, key):    return {"I": orc.dpool(ints(key)[0], gin(key, 0), gin(key, 1), ints(key)[7])}

    def batchnorm(self, key):
        N, H, W, Cc = ints(key)
        o, xh, a, r = orc.batchnorm(gin(key, 0).reshape(N, H * W, Cc), gin(key, 1), gin(key, 2))
        return {"O": o, "XH": xh, "scr": np.concatenate([r, a, np.zeros(Cc, np.float32)])}

    def dbatchnorm(self, key):
        N, H, W, Cc, tr = ints(key)
        dX, dW, dB = orc.dbatchnorm(gin(key, 0).reshape(N, H * W, Cc), gin(key, 1).reshape(N, H * W, Cc), gin(key, 2),
                                    gin(key, 5)[:Cc], gin(key, 3), gin(key, 4), bool(tr))
        return {"dX": dX, "dW": dW, "dB": dB}

    def optim(self, key, kind):
        a = [orc.f32(gin(key, k)).copy() for k in range(4 if kind != "sgd" else 3)]
        P = orc._p; f = flts(key); n = a[0].size
        if kind == "sgd":   orc.lib().orc_sgd(P(a[0]), P(a[1]), P(a[2]), ints(key)[0], f[0], f[1], n)
        elif kind == "adam": orc.lib().orc_adam(P(a[0]), P(a[1]), P(a[2]), P(a[3]), f[0], f[1], f[2], n)
        else:               orc.lib().orc_adamw(P(a[0]), P(a[1]), P(a[2]), P(a[3]), f[0], f[1], f[2], f[3], n)
        return dict(zip(["G", "DG", "M", "V"], a))


class Gpu:
    name = "cuda"

    def __init__(self):
        from gpu_util import lib, dev, zeros, ptr, host, ok
        self.L, self.dev, self.zeros, self.ptr, self.host, self.ok = lib(), dev, zeros, ptr, host, ok

    def gemm(self, key):
        v, tA, tB, M, N, K, Cc = ints(key); al, be = flts(key)
        o = self.dev(gin(key, 2))
        self.ok(self.L.t4k_gemm(self.ptr(self.dev(gin(key, 0))), self.ptr(self.dev(gin(key, 1))), self.ptr(o), al, be, tA, tB, M, N, K, Cc, 1, 0, 0, 0, None))
        return {"O": self.host(o)}

    def map(self, key):
        d = self.dev(gin(key, 0)); self.ok(self.L.t4k_map(ints(key)[0], self.ptr(d), flts(key)[0], d.numel(), None)); return {"A": self.host(d)}

    def ts_op(self, key):
        a, o = self.dev(gin(key, 0)), self.zeros(gin(key, 0).size)
        self.ok(self.L.t4k_ts_op(ints(key)[0], self.ptr(a), flts(key)[0], self.ptr(o), a.numel(), None)); return {"O": self.host(o)}

    def tt_op(self, key):
        a, b, o = self.dev(gin(key, 0)), self.dev(gin(key, 1)), self.zeros(gin(key, 0).size)
        self.ok(self.L.t4k_tt_op(ints(key)[0], self.ptr(a), self.ptr(b), self.ptr(o), a.numel(), 1, 1, None)); return {"O": self.host(o)}

    def transpose(self, key):
        H, W, Cc = ints(key); a, o = self.dev(gin(key, 0)), self.zeros(H * W * Cc)
        self.ok(self.L.t4k_transpose(self.ptr(a), self.ptr(o), 1, H, W, Cc, None)); return {"T": self.host(o)}

    def sum(self, key):
        a, o = self.dev(gin(key, 0)), self.zeros(1); self.ok(self.L.t4k_sum(self.ptr(a), a.numel(), self.ptr(o), None)); return {"v": self.host(o)}

    def nvar(self, key):
        a, o = self.dev(gin(key, 0)), self.zeros(1); self.ok(self.L.t4k_nvar(self.ptr(a), flts(key)[0], a.numel(), self.ptr(o), None)); return {"v": self.host(o)}

    def minmax(self, key):
        a, o = self.dev(gin(key, 0)), self.zeros(1); self.ok(self.L.t4k_minmax(self.ptr(a), a.numel(), ints(key)[0], self.ptr(o), None)); return {"v": self.host(o)}

    def dot(self, key):
        K, Cc = ints(key); al, be = flts(key); a, b, o = self.dev(gin(key, 0)), self.dev(gin(key, 1)), self.dev(gin(key, 2))
        self.ok(self.L.t4k_dot(self.ptr(a), self.ptr(b), self.ptr(o), al, be, K, Cc, 1, 1, None)); return {"O": self.host(o)}

    def bce(self, key):     # reference kernel returns Σ[...]; the C-ABI returns the finished loss -Σ/N → undo with N=1
        t, o, l = self.dev(gin(key, 0)), self.dev(gin(key, 1)), self.zeros(1)
        self.ok(self.L.t4k_loss(1, self.ptr(o), self.ptr(t), t.numel(), 1, self.ptr(l), None)); return {"v": -self.host(l)}

    def bias(self, key):
        N, E0 = ints(key); b, y = self.dev(gin(key, 0)), self.dev(gin(key, 1))
        self.ok(self.L.t4k_bias(self.ptr(b), self.ptr(y), N, E0, None)); return {"Y": self.host(y)}

    def dlinear_db(self, key):
        N, E0 = ints(key); dy, db = self.dev(gin(key, 0)), self.dev(gin(key, 1))
        self.ok(self.L.t4k_dbias(self.ptr(dy), self.ptr(db), N, E0, None)); return {"dB": self.host(db)}

    def activate(self, key):
        i, o, f = self.dev(gin(key, 0)), self.zeros(gin(key, 0).size), self.dev(gin(key, 2))
        self.ok(self.L.t4k_activate_fwd(ints(key)[0], self.ptr(i), self.ptr(o), self.ptr(f), flts(key)[0], i.numel(), None))
        return {"O": self.host(o), "F": self.host(f)}

    def softmax(self, key):
        N, Cc = ints(key); i, o = self.dev(gin(key, 0)), self.zeros(N * Cc)
        self.ok(self.L.t4k_softmax_fwd(self.ptr(i), self.ptr(o), N, Cc, None)); return {"O": self.host(o)}

    def conv2d(self, key):
        d = ints(key); N, H1, W1, C1, H0, W0, C0, K, S, P = d
        i, f, b, o = self.dev(gin(key, 0)), self.dev(gin(key, 1)), self.dev(gin(key, 2)), self.zeros(N * H0 * W0 * C0)
        self.ok(self.L.t4k_conv2d_fwd(self.ptr(i), self.ptr(f), self.ptr(b), self.ptr(o), *d, None)); return {"O": self.host(o)}

    def dconv2d(self, key):
        d = ints(key)
        i, do, f, dx, df, db = [self.dev(gin(key, k)) for k in range(6)]
        self.ok(self.L.t4k_conv2d_bwd(self.ptr(i), self.ptr(do), self.ptr(f), self.ptr(dx), self.ptr(df), self.ptr(db), *d, None))
        return {"dX": self.host(dx), "dF": self.host(df), "dB": self.host(db)}

    def pool(self, key):
        d = ints(key); i, o = self.dev(gin(key, 0)), self.zeros(gin(key, 1).size)
        self.ok(self.L.t4k_pool_fwd(d[0], self.ptr(i), self.ptr(o), *d[1:], None)); return {"O": self.host(o)}

    def dpool(self, key):
        d = ints(key); i, do = self.dev(gin(key, 0)), self.dev(gin(key, 1))
        self.ok(self.L.t4k_pool_bwd(d[0], self.ptr(i), self.ptr(do), *d[1:], None)); return {"I": self.host(i)}

    def batchnorm(self, key):
        N, H, W, Cc = ints(key)
        i, g, b = self.dev(gin(key, 0)), self.dev(gin(key, 1)), self.dev(gin(key, 2))
        o, xh, scr = self.zeros(i.numel()), self.zeros(i.numel()), self.zeros(3 * Cc)
        self.ok(self.L.t4k_batchnorm_fwd(self.ptr(i), self.ptr(o), self.ptr(xh), self.ptr(g), self.ptr(b), self.ptr(scr), N, H * W, Cc, None))
        return {"O": self.host(o), "XH": self.host(xh), "scr": self.host(scr)}

    def dbatchnorm(self, key):
        N, H, W, Cc, tr = ints(key)
        dy, xh, g, dw, db, scr = [self.dev(gin(key, k)) for k in range(6)]
        dx = self.zeros(dy.numel())
        self.ok(self.L.t4k_batchnorm_bwd(self.ptr(dy), self.ptr(xh), self.ptr(dx), self.ptr(g), self.ptr(dw), self.ptr(db), self.ptr(scr), N, H * W, Cc, tr, None))
        return {"dX": self.host(dx), "dW": self.host(dw), "dB": self.host(db)}

    def optim(self, key, kind):
        a = [self.dev(gin(key, k)) for k in range(4 if kind != "sgd" else 3)]
        f = flts(key); n = a[0].numel(); p = self.ptr
        if kind == "sgd":    self.ok(self.L.t4k_sgd(p(a[0]), p(a[1]), p(a[2]), ints(key)[0], f[0], f[1], n, None))
        elif kind == "adam": self.ok(self.L.t4k_adam(p(a[0]), p(a[1]), p(a[2]), p(a[3]), f[0], f[1], f[2], n, None))
        else:                self.ok(self.L.t4k_adamw(p(a[0]), p(a[1]), p(a[2]), p(a[3]), f[0], f[1], f[2], f[3], n, None))
        return dict(zip(["G", "DG", "M", "V"], [self.host(t) for t in a]))


def run_case(be, key):
    fam = key.split("_")[0]
    exact = False
    kw = {}
    if fam.startswith("gemm"):   out = be.gemm(key)
    elif fam == "map":
        out = be.map(key); exact = ints(key)[0] in EXACT_MAP; kw = dict(rtol=1e-5, atol=1e-6)
    elif fam == "ts":            out = be.ts_op(key); exact = True
    elif fam == "tt":            out = be.tt_op(key); exact = True
    elif fam == "transpose":     out = be.transpose(key); exact = True
    elif fam == "sum":           out = be.sum(key); kw = dict(rtol=1e-5, atol=1e-3)
    elif fam == "nvar":          out = be.nvar(key); kw = dict(rtol=1e-5)
    elif fam in ("max", "min"):  out = be.minmax(key); exact = True
    elif fam == "dot":           out = be.dot(key); kw = dict(rtol=1e-5)
    elif fam == "bce":           out = be.bce(key); kw = dict(rtol=1e-5)
    elif fam == "bias":          out = be.bias(key); exact = True
    elif fam == "dlinear":       out = be.dlinear_db(key)
    elif fam == "act":
        out = be.activate(key); exact = ints(key)[0] in (orc.L_RELU, orc.L_LEAKYRL, orc.L_DROPOUT); kw = dict(rtol=1e-5, atol=1e-6)
    elif fam == "softmax":       out = be.softmax(key); kw = dict(rtol=1e-5, atol=1e-7)
    elif fam == "conv":          out = be.conv2d(key)
    elif fam == "dconv":         out = be.dconv2d(key)
    elif fam == "pool":          out = be.pool(key); exact = ints(key)[0] != orc.L_AVGPOOL
    elif fam in ("dpool", "upsample"): out = be.dpool(key); exact = ints(key)[0] != orc.L_AVGPOOL
    elif fam == "bn":            out = be.batchnorm(key)
    elif fam == "dbn":           out = be.dbatchnorm(key)
    elif fam in ("sgd0", "sgdm"): out = be.optim(key, "sgd"); kw = dict(rtol=1e-5, atol=1e-7)
    elif fam == "adam":          out = be.optim(key, "adam"); kw = dict(rtol=1e-5, atol=1e-7)
    elif fam == "adamw":         out = be.optim(key, "adamw"); kw = dict(rtol=1e-5, atol=1e-7)
    else:
        raise KeyError(key)
    checked = 0
    for nm, got in out.items():
        rk = "%s/%s" % (key, nm)
        if rk not in G.files:
            continue
        ref = G[rk]
        if key == "bn" and nm == "scr":            # [2C,3C) is backward scratch, unspecified after forward
            Cc = ints(key)[3]; got, ref = np.asarray(got).ravel()[:2 * Cc], ref[:2 * Cc]
        close(got, ref, exact=exact, what="%s[%s] %s" % (be.name, key, nm), **kw)
        checked += 1
    assert checked > 0, key


@pytest.mark.parametrize("key", KEYS)
def test_oracle_reproduces_reference_kernels(key):
    run_case(Orc(), key)


@pytest.mark.gpu
@pytest.mark.parametrize("key", KEYS)
def test_cuda_path_reproduces_reference_kernels(key):
    run_case(Gpu(), key)


# ------------------------------------------------------------------ conv-transpose (L_DCONV): the reference's kernels with swapped roles
def test_conv_transpose_oracle_is_the_reference_kernels_with_swapped_roles():
    """oracle.convt2d / dconvt2d (what t4k_dconv2d_fwd / _bwd are held to) against the outputs of the reference's OWN k_dconv2d and k_conv2d at the
    layer's configuration (4x4, stride 2, padding 1: fixtures dconv_k4 / conv_k4, produced by oracle/ref/refkern.cu from the unmodified nmath.tcu):
      forward   = k_dconv2d's dX with (output gradient := layer input)                       (src/nn/forward.cu:110)
      backward  = k_conv2d of the layer's output gradient (bias-free) and k_dconv2d's dF with (input, output gradient) := (dO, I)  (backprop.cu:137)"""
    key = "dconv_k4"
    N, H, W, C1, H0, W0, C0, K, S, P, tr = ints(key)
    big, small, F = gin(key, 0).reshape(N, H, W, C1), gin(key, 1).reshape(N, H0, W0, C0), gin(key, 2).reshape(C1, K, K, C0)                   # conv input [N,14,14,2], its output gradient [N,7,7,6], filter [2][4][4][6]
    assert orc.convt_out_dims(H0, W0, K, S, P) == (H + 1, W + 1)            # model.cpp:129-133 gives an odd input the larger of the two pre-images (15);
    # forward: layer input = `small`; no bias; the fixture's own pre-image (14) exercises the same kernel
    O = orc.convt2d(small, F, np.zeros(C1, np.float32), K, S, P, out_hw=(H, W))
    close(O, G[key + "/dX"], what="conv-transpose forward == the reference's k_dconv2d dX")
    # backward, parameters: layer input `small`, output gradient `big`
    dX, dF, dB = orc.dconvt2d(small, big, F, K, S, P, dF=gin(key, 4).reshape(C1, K, K, C0), dB=np.zeros(C1, np.float32), train=True)
    close(dF, G[key + "/dF"], what="conv-transpose dF == the reference's k_dconv2d dF")
    close(dB, big.reshape(-1, C1).astype(np.float64).sum(0), rtol=1e-5, what="dB = pixel sums of dO")
    # backward, input gradient: the reference's k_conv2d on the same input (fixture conv_k4 carries a bias: removed)
    ck = "conv_k4"
    want = G[ck + "/O"].astype(np.float64).reshape(-1, C0) - gin(ck, 2).astype(np.float64)
    got, _, _ = orc.dconvt2d(np.zeros((N, H0, W0, C0), np.float32), gin(ck, 0).reshape(N, H, W, C1), gin(ck, 1).reshape(C1, K, K, C0), K, S, P, train=False)
    close(got, want, rtol=1e-5, what="conv-transpose dX == the reference's k_conv2d (bias removed)")
