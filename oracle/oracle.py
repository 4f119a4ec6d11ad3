"""
oracle.py — ctypes front-end of oracle/libt4oracle.so + a restatement of the reference's
Tensor/Model host logic (layer construction, forward/backprop order, optimizer loop).

TEST INFRASTRUCTURE ONLY.  Import allowed from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs; never from tensorforth_b200/.

All reference citations are file:line relative to /root/reference.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libt4oracle.so")

# enum mirrors (src/t4math.h:25-56, src/nn/ntypes.h:16-43)
ABS, NEG, EXP, LN, LOG, TANH, RELU, SIGM, SQRT, RCP, SAT, IDEN, FILL, GFILL, SCALE, POW, \
    ADD, SUB, MUL, DIV, MOD, MAX, MIN = range(23)
(L_NONE, L_CONV, L_LINEAR, L_FLATTEN, L_RELU, L_TANH, L_SIGMOID, L_SELU, L_LEAKYRL, L_ELU,
 L_DROPOUT, L_SOFTMAX, L_LOGSMAX, L_AVGPOOL, L_MAXPOOL, L_MINPOOL, L_BATCHNM, L_USAMPLE,
 L_DCONV) = range(19)
LOSS_MSE, LOSS_BCE, LOSS_CE, LOSS_NLL = range(4)


def build():
    """compile the C oracle (gcc) if missing or stale"""
    src = os.path.join(_HERE, "t4_oracle.c")
    if (not os.path.exists(_SO)) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "libt4oracle.so"])


_lib = None
_fp = C.POINTER(C.c_float)
_ip = C.POINTER(C.c_int)


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        f, i, l = C.c_float, C.c_int, C.c_long
        sig = {
            "orc_gemm": (None, [_fp, _fp, _fp, f, f, i, i, i, i, i, i]),
            "orc_gemm_f64acc": (None, [_fp, _fp, _fp, f, f, i, i, i, i]),
            "orc_map": (None, [i, _fp, f, l]),
            "orc_ts_op": (None, [i, _fp, f, _fp, l]),
            "orc_tt_op": (None, [i, _fp, _fp, _fp, l]),
            "orc_copy": (None, [_fp, _fp, l]),
            "orc_transpose": (None, [_fp, _fp, i, i, i]),
            "orc_identity": (None, [_fp, i, i, i]),
            "orc_sum": (f, [_fp, l]),
            "orc_nvar": (f, [_fp, f, l]),
            "orc_max": (f, [_fp, l, i]),
            "orc_dot": (None, [_fp, _fp, _fp, f, f, i, i]),
            "orc_bce_sum": (f, [_fp, _fp, l]),
            "orc_avg": (f, [_fp, l]),
            "orc_std": (f, [_fp, l]),
            "orc_norm": (f, [_fp, l]),
            "orc_loss": (f, [i, _fp, _fp, l, i]),
            "orc_bias": (None, [_fp, _fp, i, i]),
            "orc_dlinear_db": (None, [_fp, _fp, i, i]),
            "orc_activate": (None, [i, _fp, _fp, _fp, f, l]),
            "orc_softmax": (None, [_fp, _fp, i, i]),
            "orc_logsoftmax": (None, [_fp, _fp, i, i]),
            "orc_conv2d": (None, [_fp, _fp, _fp, _fp] + [i] * 10),
            "orc_dconv2d": (None, [_fp] * 6 + [i] * 11),
            "orc_pool": (None, [i, _fp, _fp] + [i] * 7),
            "orc_dpool": (None, [i, _fp, _fp] + [i] * 7),
            "orc_batchnorm": (None, [_fp] * 7 + [i] * 3),
            "orc_dbatchnorm": (None, [_fp] * 9 + [i] * 4),
            "orc_sgd": (None, [_fp, _fp, _fp, i, f, f, l]),
            "orc_adam": (None, [_fp, _fp, _fp, _fp, f, f, f, l]),
            "orc_adamw": (None, [_fp, _fp, _fp, _fp, f, f, f, f, l]),
            "orc_onehot": (None, [_ip, _fp, i, i]),
            "orc_hit": (i, [_fp, _fp, i, i]),
        }
        for name, (res, args) in sig.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def _p(a):
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"], (a.dtype, a.flags)
    return a.ctypes.data_as(_fp)


def f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


# ---------------------------------------------------------------------------------------
# kernel-level wrappers (functional style: inputs are not modified unless stated)
# ---------------------------------------------------------------------------------------
def gemm(A, B, O=None, alpha=1.0, beta=0.0, tA=False, tB=False, M=None, N=None, K=None, C=1):
    """Tensor::linear / gemm3 per sample — src/mu/tensor.cu:80-87,162-180"""
    A, B = f32(A), f32(B)
    if M is None:
        a2 = A.reshape(A.shape[0], -1) if C == 1 else None
        M, K = (A.shape[1], A.shape[0]) if tA else (A.shape[0], A.shape[1])
        N = B.shape[0] if tB else B.shape[1]
    O = np.zeros((M, N, C) if C > 1 else (M, N), np.float32) if O is None else f32(O).copy()
    lib().orc_gemm(_p(A), _p(B), _p(O), alpha, beta, int(tA), int(tB), M, N, K, C)
    return O


def map_(op, A, v=0.0):
    A = f32(A).copy()
    lib().orc_map(op, _p(A), float(v), A.size)
    return A


def ts_op(op, A, v):
    A = f32(A)
    O = np.empty_like(A)
    lib().orc_ts_op(op, _p(A), float(v), _p(O), A.size)
    return O


def tt_op(op, A, B):
    """Tensor::ten_op incl. the N-broadcast of src/mu/tensor.cu:39-46 (N is axis 0)"""
    A, B = f32(A), f32(B)
    if A.size == B.size:
        O = np.empty_like(A)
        lib().orc_tt_op(op, _p(A), _p(B), _p(O), A.size)
        return O
    big, small, a_small = (B, A, True) if A.size < B.size else (A, B, False)
    n = big.shape[0]
    O = np.empty_like(big)
    for k in range(n):
        s = f32(big[k])
        o = np.empty_like(s)
        if a_small:
            lib().orc_tt_op(op, _p(f32(small.reshape(s.shape))), _p(s), _p(o), s.size)
        else:
            lib().orc_tt_op(op, _p(s), _p(f32(small.reshape(s.shape))), _p(o), s.size)
        O[k] = o
    return O


def transpose(A, C_=1):
    A = f32(A)
    H, W = A.shape[0], A.shape[1]
    T = np.empty((W, H) + A.shape[2:], np.float32)
    lib().orc_transpose(_p(A), _p(T), H, W, C_)
    return T


def identity(H, W, C_=1):
    T = np.empty((H, W, C_), np.float32)
    lib().orc_identity(_p(T), H, W, C_)
    return T


def tsum(A):  return float(lib().orc_sum(_p(f32(A)), A.size))
def avg(A):   return float(lib().orc_avg(_p(f32(A)), A.size))
def std(A):   return float(lib().orc_std(_p(f32(A)), A.size))
def norm(A):  return float(lib().orc_norm(_p(f32(A)), A.size))
def tmax(A):  return float(lib().orc_max(_p(f32(A)), A.size, 1))
def tmin(A):  return float(lib().orc_max(_p(f32(A)), A.size, 0))


def dot(A, B, O=None, alpha=1.0, beta=0.0, C_=1):
    A, B = f32(A), f32(B)
    K = A.size // C_
    O = np.zeros(C_, np.float32) if O is None else f32(O).copy()
    lib().orc_dot(_p(A), _p(B), _p(O), alpha, beta, K, C_)
    return O


def loss(op, out, tgt, N):
    """Tensor::loss on a copy — src/mu/tensor.cu:289-325"""
    o = f32(out).copy()
    t = f32(tgt)
    return float(lib().orc_loss(op, _p(o), _p(t), o.size, N))


def activate(layer, I, alpha=0.0, mask=None):
    I = f32(I)
    O = np.empty_like(I)
    F = np.zeros_like(I) if mask is None else f32(mask).copy()
    lib().orc_activate(layer, _p(I), _p(O), _p(F), float(alpha), I.size)
    return O, F


def softmax(I, N):
    I = f32(I); O = np.empty_like(I)
    lib().orc_softmax(_p(I), _p(O), N, I.size // N)
    return O


def logsoftmax(I, N):
    I = f32(I); O = np.empty_like(I)
    lib().orc_logsoftmax(_p(I), _p(O), N, I.size // N)
    return O


def conv_out_dims(H1, W1, K, S, P):
    """Model::_iconv — src/nn/model.cpp:136-137 (W0 computed from H1, sic)"""
    H0 = (H1 - K + 2 * P) // S + 1
    W0 = (H1 - K + 2 * P) // S + 1
    return H0, W0


def conv2d(I, F, B, K, S, P):
    I, F, B = f32(I), f32(F), f32(B)
    N, H1, W1, C1 = I.shape
    C0 = F.shape[3]
    H0, W0 = conv_out_dims(H1, W1, K, S, P)
    O = np.empty((N, H0, W0, C0), np.float32)
    lib().orc_conv2d(_p(I), _p(F), _p(B), _p(O), N, H1, W1, C1, H0, W0, C0, K, S, P)
    return O


def dconv2d(I, dO, F, K, S, P, dF=None, dB=None, train=True):
    I, dO, F = f32(I), f32(dO), f32(F)
    N, H1, W1, C1 = I.shape
    _, H0, W0, C0 = dO.shape
    dX = np.empty_like(I)
    dF = np.zeros_like(F) if dF is None else f32(dF).copy()
    dB = np.zeros(C0, np.float32) if dB is None else f32(dB).copy()
    lib().orc_dconv2d(_p(I), _p(dO), _p(F), _p(dX), _p(dF), _p(dB),
                      N, H1, W1, C1, H0, W0, C0, K, S, P, int(train))
    return dX, dF, dB


def convt_out_dims(H1, W1, K, S, P):
    """Model::_iconv, transposed branch (src/nn/model.cpp:129-133)"""
    P0 = (H1 + 2 * P - K) % S
    return (H1 - 1) * S - 2 * P + K + P0, (W1 - 1) * S - 2 * P + K + P0


def convt2d(I, F, B, K, S, P, out_hw=None):
    """conv-transpose forward as the reference wires L_DCONV (src/nn/forward.cu:110 -> Model::_bconv): the input-gradient half of k_dconv2d
    (orc_dconv2d, flipped taps) applied to the layer input, on the geometry of the convolution (C0 -> C1) that maps the large image onto the small
    one; F [C0][K][K][C1]; then k_bias per output channel."""
    I, F, B = f32(I), f32(F), f32(B)
    N, H1, W1, C1 = I.shape
    C0 = F.shape[0]
    H0, W0 = out_hw if out_hw else convt_out_dims(H1, W1, K, S, P)   # out_hw: any large size the convolution maps onto (H1, W1)
    assert conv_out_dims(H0, W0, K, S, P) == (H1, W1)
    O, _, _ = dconv2d(np.zeros((N, H0, W0, C0), np.float32), I, F, K, S, P, train=False)
    lib().orc_bias(_p(B), _p(O), N * H0 * W0, C0)
    return O


def dconvt2d(I, dO, F, K, S, P, dF=None, dB=None, train=True):
    """conv-transpose backward (src/nn/backprop.cu:137 -> Model::_fconv): dX = k_conv2d(dO, F) without bias; train: dF += the filter-gradient half
    of k_dconv2d on (input, output gradient) = (dO, I), dB[c0] += sum of dO over the pixels"""
    I, dO, F = f32(I), f32(dO), f32(F)
    C0, C1 = F.shape[0], F.shape[3]
    dF = np.zeros_like(F) if dF is None else f32(dF).copy()
    dB = np.zeros(C0, np.float32) if dB is None else f32(dB).copy()
    if train:
        _, dF, _ = dconv2d(dO, I, F, K, S, P, dF=dF, dB=np.zeros(C1, np.float32), train=True)
        dB = (dB + dO.reshape(-1, C0).sum(axis=0, dtype=np.float64)).astype(np.float32)
    dX = conv2d(dO, F, np.zeros(C1, np.float32), K, S, P)
    return dX, dF, dB


def pool(layer, I, K):
    I = f32(I)
    N, H1, W1, Cc = I.shape
    H0, W0 = (H1 + K - 1) // K, (W1 + K - 1) // K         # src/nn/model.cpp:267-268
    O = np.empty((N, H0, W0, Cc), np.float32)
    lib().orc_pool(layer, _p(I), _p(O), N, H1, W1, H0, W0, Cc, K)
    return O


def dpool(layer, I, dO, K):
    """in-place semantics of k_dpool on a copy of the forward input"""
    I = f32(I).copy(); dO = f32(dO)
    N, H1, W1, Cc = I.shape
    _, H0, W0, _ = dO.shape
    lib().orc_dpool(layer, _p(I), _p(dO), N, H1, W1, H0, W0, Cc, K)
    return I


def batchnorm(I, gamma, beta):
    I = f32(I)
    Cc = I.shape[-1]; N = I.shape[0]; HW = I.size // (N * Cc)
    O = np.empty_like(I); XH = np.empty_like(I)
    avg_ = np.empty(Cc, np.float32); rvar = np.empty(Cc, np.float32)
    lib().orc_batchnorm(_p(I), _p(O), _p(XH), _p(f32(gamma)), _p(f32(beta)), _p(avg_), _p(rvar), N, HW, Cc)
    return O, XH, avg_, rvar


def dbatchnorm(dO, XH, gamma, rvar, dW, dB, train=True):
    dO, XH = f32(dO), f32(XH)
    Cc = dO.shape[-1]; N = dO.shape[0]; HW = dO.size // (N * Cc)
    dX = np.empty_like(dO)
    dW = f32(dW).copy(); dB = f32(dB).copy()
    s1 = np.empty(Cc, np.float32); s2 = np.empty(Cc, np.float32)
    lib().orc_dbatchnorm(_p(dO), _p(XH), _p(dX), _p(f32(gamma)), _p(dW), _p(dB), _p(f32(rvar)),
                         _p(s1), _p(s2), N, HW, Cc, int(train))
    return dX, dW, dB


def onehot(labels, E):
    lab = np.ascontiguousarray(labels, dtype=np.int32)
    hot = np.empty((lab.size, E), np.float32)
    lib().orc_onehot(lab.ctypes.data_as(_ip), _p(hot), lab.size, E)
    return hot


def dataset_normalize(mean, scale):
    """Dataset::normalize (src/mu/dataset.cu:33-41): stores (mean, 1/scale); scale == 0 -> error, 1.0"""
    return np.float32(mean), (np.float32(1.0) if abs(scale) < 1e-6 else np.float32(1.0) / np.float32(scale))


def dataset_load(u8, mean=0.0, scale=1.0 / 256.0):
    """Dataset::_load (src/mu/dataset.cu:139-143): d[i] = (I2D((int)u8[i]) - _mean) * _scale, two FP32 roundings.
    `mean`/`scale` are the STORED values (dataset.h:35-36 defaults: 0, 1/256; see dataset_normalize)."""
    x = np.ascontiguousarray(u8, dtype=np.uint8).astype(np.int32).astype(np.float32)
    return ((x - np.float32(mean)).astype(np.float32) * np.float32(scale)).astype(np.float32)


def hit(out, hot):
    out, hot = f32(out), f32(hot)
    N = out.shape[0]
    return int(lib().orc_hit(_p(out), _p(hot), N, out.size // N))


# ---------------------------------------------------------------------------------------
# Model-level restatement (src/nn/model.cpp, forward.cu, backprop.cu, gradient.cu)
# ---------------------------------------------------------------------------------------
class Layer:
    """One Model layer tensor: holds that layer's INPUT (src/nn/model.cpp:153-157)."""
    def __init__(self, shape):
        self.data = np.zeros(shape, np.float32)      # [N,H,W,C]
        self.fn = L_NONE
        self.w = self.b = self.dw = self.db = self.ex = None     # grad[0..4]
        self.m = [None, None, None, None]            # mtum[0..3] = m_w, m_b, v_w, v_b
        self.xparm = 0.0
        self.K = self.S = self.P = 0
        self.bn_scratch = None


class OracleModel:
    """Restates Model::add/forward/backprop/sgd/adam (src/nn/model.cpp:83-310,
    forward.cu:29-113, backprop.cu:40-140, gradient.cu:64-169)."""

    def __init__(self, N, H, W, Cc, seed=0):
        self.layers = [Layer((N, H, W, Cc))]
        self.train = True
        self.rng = np.random.default_rng(seed)
        self._iter = 0
        self.epoch = 0
        self._opt_alloc = None

    # ---- construction (src/nn/model.cpp:122-310; vocabulary src/vm/netvm.cpp:20-133) ----
    def _rand(self, shape, scale):
        # Model::RAND: scale*2*(-0.5 + U(0,1]) = [-scale, scale)  (src/nn/model.cpp:74-79)
        return ((self.rng.random(shape, dtype=np.float32) - 0.5) * (2.0 * scale)).astype(np.float32)

    def add(self, fn, n=0, bias=0.0, opt=None):
        t = self.layers[-1]
        N, H, W, Cc = t.data.shape
        t.fn = fn
        if fn in (L_CONV,):
            K = opt[0] if opt else 3
            S = opt[1] if opt else 1
            P = opt[2] if (opt and K > 1 and opt[2]) else (K - 1) // 2
            t.K, t.S, t.P = K, S, P
            t.xparm = bias
            k = np.float32(np.sqrt(6.0 / (K * K * Cc)))
            t.w = self._rand((Cc, K, K, n), k)
            t.b = self._rand((n,), bias)
            t.dw = np.zeros_like(t.w); t.db = np.zeros_like(t.b)
            t.ex = np.zeros_like(t.data)
            H0, W0 = conv_out_dims(H, W, K, S, P)
            self.layers.append(Layer((N, H0, W0, n)))
        elif fn == L_DCONV:
            K = opt[0] if opt else 4
            S = opt[1] if opt else 2
            P = opt[2] if (opt and K > 1 and opt[2]) else (K - 1) // 2
            t.K, t.S, t.P = K, S, P
            t.xparm = bias
            k = np.float32(np.sqrt(6.0 / (K * K * Cc)))
            t.w = self._rand((n, K, K, Cc), k)              # [C0][K][K][C1]: see convt2d
            t.b = self._rand((n,), bias)
            t.dw = np.zeros_like(t.w); t.db = np.zeros_like(t.b)
            t.ex = np.zeros_like(t.data)
            H0, W0 = convt_out_dims(H, W, K, S, P)
            self.layers.append(Layer((N, H0, W0, n)))
        elif fn == L_LINEAR:
            E1 = H * W * Cc
            k = np.float32(np.sqrt(1.0 / (n + E1)))
            t.xparm = bias
            t.w = self._rand((n, E1), k)
            t.b = self._rand((n,), bias)
            t.dw = np.zeros_like(t.w); t.db = np.zeros_like(t.b)
            self.layers.append(Layer((N, 1, n, 1)))
        elif fn == L_FLATTEN:
            self.layers.append(Layer((N, 1, H * W * Cc, 1)))
        elif fn in (L_RELU, L_TANH, L_SIGMOID, L_SELU, L_LEAKYRL, L_ELU, L_DROPOUT):
            t.ex = np.zeros_like(t.data); t.xparm = bias
            self.layers.append(Layer((N, H, W, Cc)))
        elif fn in (L_SOFTMAX, L_LOGSMAX):
            self.layers.append(Layer((N, H, W, Cc)))
        elif fn in (L_AVGPOOL, L_MAXPOOL, L_MINPOOL):
            t.K = n
            self.layers.append(Layer((N, (H + n - 1) // n, (W + n - 1) // n, Cc)))
        elif fn == L_BATCHNM:
            t.w = np.ones(Cc, np.float32); t.b = np.zeros(Cc, np.float32)
            t.dw = np.zeros(Cc, np.float32); t.db = np.zeros(Cc, np.float32)   # reference leaves these uninitialised (model.cpp:281,283)
            t.ex = np.zeros_like(t.data); t.xparm = bias
            t.bn_scratch = np.zeros(3 * Cc, np.float32)
            self.layers.append(Layer((N, H, W, Cc)))
        elif fn == L_USAMPLE:
            t.K = n
            self.layers.append(Layer((N, H * n, W * n, Cc)))
        else:
            raise ValueError(fn)
        return self

    # ---- forward (src/nn/forward.cu:29-113) ----
    def forward(self, x, dropout_masks=None):
        L = self.layers
        assert x.size == L[0].data.size
        L[0].data = f32(x).reshape(L[0].data.shape).copy()
        for i in range(len(L) - 1):
            t, o = L[i], L[i + 1]
            fn = t.fn
            N = t.data.shape[0]
            if fn == L_CONV:
                o.data = conv2d(t.data, t.w, t.b, t.K, t.S, t.P)
            elif fn == L_DCONV:
                o.data = convt2d(t.data, t.w, t.b, t.K, t.S, t.P)
            elif fn == L_LINEAR:
                y = gemm(t.data.reshape(N, -1), t.w, tB=True)
                lib().orc_bias(_p(t.b), _p(y), N, y.shape[1])
                o.data = y.reshape(o.data.shape)
            elif fn == L_FLATTEN:
                o.data = t.data.reshape(o.data.shape).copy()
            elif fn in (L_RELU, L_TANH, L_SIGMOID, L_SELU, L_LEAKYRL, L_ELU):
                o.data, t.ex = activate(fn, t.data, t.xparm)
            elif fn == L_DROPOUT:
                u = dropout_masks[i] if dropout_masks is not None else 1.0 - self.rng.random(t.data.shape, dtype=np.float32)
                o.data, t.ex = activate(fn, t.data, t.xparm, mask=u)
            elif fn == L_SOFTMAX:
                o.data = softmax(t.data, N).reshape(o.data.shape)
            elif fn == L_LOGSMAX:
                o.data = logsoftmax(t.data, N).reshape(o.data.shape)
            elif fn in (L_AVGPOOL, L_MAXPOOL, L_MINPOOL):
                o.data = pool(fn, t.data, t.K)
            elif fn == L_BATCHNM:
                o.data, t.ex, a, r = batchnorm(t.data, t.w, t.b)
                Cc = a.size
                t.bn_scratch[:Cc] = r; t.bn_scratch[Cc:2 * Cc] = a
            elif fn == L_USAMPLE:
                z = np.zeros(o.data.shape, np.float32)
                o.data = dpool(L_USAMPLE, z, t.data, t.K)
            else:
                raise ValueError(fn)
        return self

    def output(self):
        return self.layers[-1].data

    def loss(self, op, tgt):
        """Model::loss — src/nn/loss.cpp:119-136 (non-destructive)"""
        out = self.layers[-1].data
        return loss(op, out, f32(tgt).reshape(out.shape), out.shape[0])

    # ---- backprop (src/nn/backprop.cu:40-140) ----
    def backprop(self, tgt):
        L = self.layers
        out = L[-1]
        tgt = f32(tgt).reshape(out.data.shape)
        fn_last = L[-2].fn
        if fn_last in (L_LINEAR, L_SIGMOID, L_SOFTMAX, L_LOGSMAX):        # _bprep :97-103
            out.data = tt_op(SUB, out.data, tgt)
        else:
            out.data = tgt.copy()
        for j, i in enumerate(range(len(L) - 2, -1, -1)):
            t, o = L[i], L[i + 1]
            fn = t.fn
            N = t.data.shape[0]
            if fn == L_CONV:
                dX, t.dw, t.db = dconv2d(t.data, o.data, t.w, t.K, t.S, t.P, t.dw, t.db, self.train)
                t.ex = dX; t.data = dX.copy()
            elif fn == L_DCONV:
                dX, t.dw, t.db = dconvt2d(t.data, o.data, t.w, t.K, t.S, t.P, t.dw, t.db, self.train)
                t.ex = dX; t.data = dX.copy()
            elif fn == L_LINEAR:
                if j == 0:                               # last layer linear + MSE :119-121
                    t.data = o.data.reshape(t.data.shape).copy()
                else:
                    dY = o.data.reshape(N, -1)
                    X = t.data.reshape(N, -1)
                    if self.train:
                        lib().orc_dlinear_db(_p(f32(dY)), _p(t.db), N, dY.shape[1])
                        t.dw = gemm(dY, X, O=t.dw, alpha=1.0, beta=1.0, tA=True)
                    t.data = gemm(dY, t.w).reshape(t.data.shape)
            elif fn in (L_FLATTEN, L_SIGMOID, L_SOFTMAX, L_LOGSMAX):       # pass-through :122,:129-131
                t.data = o.data.reshape(t.data.shape).copy()
            elif fn in (L_RELU, L_TANH, L_SELU, L_LEAKYRL, L_ELU, L_DROPOUT):
                t.data = tt_op(MUL, o.data, t.ex)
            elif fn in (L_AVGPOOL, L_MAXPOOL, L_MINPOOL):
                t.data = dpool(fn, t.data, o.data, t.K)
            elif fn == L_BATCHNM:
                Cc = t.w.size
                t.data, t.dw, t.db = dbatchnorm(o.data, t.ex, t.w, t.bn_scratch[:Cc], t.dw, t.db, self.train)
            elif fn == L_USAMPLE:
                t.data = pool(L_USAMPLE, o.data, t.K)
            else:
                raise ValueError(fn)
        return self

    # ---- optimizers (src/nn/gradient.cu:64-169) ----
    def _params(self):
        for t in self.layers[:-1]:
            if t.w is not None and t.dw is not None:
                # Nw = parameter tensor's N(): conv filter T4(C1,K,K,C0) → C1; linear T4(1,..) → 1; VEC → 1
                yield t, "w", (t.w.shape[0] if t.fn == L_CONV else t.w.shape[3] if t.fn == L_DCONV else 1)   # dconv2d: the reference's T4(C1,K,K,C0) -> C1
                yield t, "b", 1

    def _step(self, kind, lr, b1, b2, wd=0.0):
        first = (self._iter == 0 and self.epoch == 0)
        self._iter += 1
        if first:
            for t, nm, _ in self._params():
                k = 0 if nm == "w" else 1
                p = getattr(t, nm)
                t.m[k] = np.zeros_like(p); t.m[k + 2] = np.zeros_like(p)
        if not self.train:
            return self
        for t, nm, Nw in self._params():
            k = 0 if nm == "w" else 1
            g, dg = getattr(t, nm), getattr(t, "d" + nm)
            if kind == "sgd":
                lib().orc_sgd(_p(g), _p(dg), _p(t.m[k]), Nw, lr, b1, g.size)
            elif kind == "adam":
                lib().orc_adam(_p(g), _p(dg), _p(t.m[k]), _p(t.m[k + 2]), lr, b1, b2, g.size)
            else:
                lib().orc_adamw(_p(g), _p(dg), _p(t.m[k]), _p(t.m[k + 2]), lr, b1, b2, wd, g.size)
        return self

    def sgd(self, lr, b=0.9):
        # Model::sgd: momentum forced to 0 on the very first call (_iter==0) — gradient.cu:139
        return self._step("sgd", lr, b if self._iter else 0.0, 0.0)

    def adam(self, lr, b1=0.9, b2=0.999):
        return self._step("adam", lr, b1, b2)

    def adamw(self, lr, wd=0.001, b1=0.9, b2=0.999):
        return self._step("adamw", lr, b1, b2, wd)
