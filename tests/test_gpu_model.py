"""
-m gpu Model-level parity: the host mirror (libt4host.so: t4::Tensor / t4::Model) driven exactly as
the reference's example scripts drive the VM, checked against (a) the scripts' own `verify`
numbers and (b) the oracle's Model restatement on the same injected weights (the
reference's RNG is time-seeded, so parity tests inject parameters with nn.w= / nn.b=, as
examples/t4_30b.4th:11-20 does).
"""
import ctypes as C
import numpy as np
import pytest
import torch

from oracle import oracle as orc
from tensorforth_b200 import lib as t4
from tensorforth_b200 import host as th
from gpu_util import assert_close

pytestmark = pytest.mark.gpu
A4 = 6e-5            # half-ulp of the reference's %+.4f print


def close4(a, b, atol=A4):
    np.testing.assert_allclose(np.asarray(a, np.float64).ravel(), np.asarray(b, np.float64).ravel(), rtol=0, atol=atol)


# ------------------------------------------------------------------ examples/t4_20a.4th (tensor words)
def test_t4_20a_tensor_words():
    A = th.Tensor.matrix(2, 3, [1, 2, 3, 4, 5, 6])
    B = th.Tensor.matrix(3, 2).ones()
    close4((A @ B).numpy(), [[6, 6], [15, 15]], 0)                           # :2-9
    R = th.Tensor.matrix(512, 1024).rand()
    O = R @ th.Tensor.matrix(1024, 256).ones()
    O /= 1024.0                                                               # :12-16
    r = R.numpy().astype(np.float64).mean(1)
    assert_close(O.numpy(), np.repeat(r[:, None], 256, 1), rtol=1e-5)
    X = th.Tensor.matrix(2, 3, [1, 2, 3, 4, 5, 6]); Y = th.Tensor.matrix(2, 3).ones()
    X += Y; close4(X.numpy(), [[2, 3, 4], [5, 6, 7]], 0)                      # :45-51
    X -= Y; X -= Y; close4(X.numpy(), [[0, 1, 2], [3, 4, 5]], 0)              # :53-55
    P = th.Tensor.matrix(2, 3, [1, 2, 3, 0, 4, 5]) @ th.Tensor.matrix(3, 2).ones()
    close4(P.numpy(), [[6, 6], [9, 9]], 0)                                    # :57-62
    Hh = th.Tensor.matrix(2, 2).ones(); Hh *= 0.5; P *= Hh
    close4(P.numpy(), [[3, 3], [4.5, 4.5]], 0)                                # :64-68
    with pytest.raises(t4.T4KError):
        th.Tensor.matrix(2, 3) @ th.Tensor.matrix(2, 3)                       # "A.W != B.H dim?"


def test_tensor_reductions_and_words():
    a = np.random.default_rng(3).random(5000, dtype=np.float32) - 0.5
    T = th.Tensor.vector(a.size, a)
    assert_close(T.sum(), orc.tsum(a), rtol=1e-4, atol=1e-3)
    assert_close(T.avg(), orc.avg(a), rtol=1e-4, atol=1e-6)
    assert_close(T.std(), orc.std(a), rtol=1e-4)
    assert_close(T.norm(), orc.norm(a), rtol=1e-4)
    assert abs(T.max() - a.max()) <= abs(a.max()) * 2e-7 and abs(T.min() - a.min()) <= abs(a.min()) * 2e-7   # SCALAR() clears bit 0
    assert_close(T.dot(T), float(np.dot(a.astype(np.float64), a.astype(np.float64))), rtol=1e-4)
    M = th.Tensor.matrix(3, 5, np.arange(15)); Mt = M.transpose()
    assert Mt.shape[1:3] == (5, 3) and np.array_equal(Mt.numpy(), np.arange(15, dtype=np.float32).reshape(3, 5).T)
    E = th.Tensor.matrix(4, 4).eye()
    assert np.array_equal(E.numpy(), np.eye(4, dtype=np.float32))
    G = th.gemm(3, M, Mt, th.Tensor.matrix(3, 3).ones(), 2.0, 0.5)
    m = np.arange(15, dtype=np.float64).reshape(3, 5)
    assert_close(G.numpy(), 2 * m @ m.T + 0.5, rtol=1e-5)


# ------------------------------------------------------------------ examples/t4_30a.4th
def test_t4_30a_linear_forward():
    nn = th.Model(1, 1, 2, 1).linear(3)
    nn.set_w(0, 0.1 * np.array([[1, 2], [3, 4], [5, 6]], np.float32)).set_b(0, [1, 2, 3])
    close4(nn.w(0).numpy(), 0.1 * np.array([1, 2, 3, 4, 5, 6]), 1e-7)
    nn.forward(th.Tensor.vector(2, [10, 20]))
    close4(nn.layer(-1).numpy(), [6, 13, 20], 1e-5)


# ------------------------------------------------------------------ examples/t4_30b.4th / t4_30c.4th
def mazur(N, hidden):
    return th.Model(N, 1, 2, 1).linear(hidden).sigmoid().linear(2).sigmoid()


def test_t4_30b_mazur_n1():
    nn = mazur(1, 3)
    nn.set_w(0, [0.15, 0.2, 0.25, 0.3, 0.2, 0.15]).set_b(0, [0.35] * 3)
    nn.set_w(2, [0.4, 0.45, 0.5, 0.55, 0.5, 0.45]).set_b(2, [0.6] * 2)
    nn.forward(th.Tensor.vector(2, [0.05, 0.1]))
    close4(nn.layer(1).numpy(), [0.3775, 0.3925, 0.3750])
    close4(nn.w(1).numpy(), [0.2413, 0.2406, 0.2414])                        # `1 nn.w` on a sigmoid layer → its filter grad[4]
    close4(nn.layer(2).numpy(), [0.5933, 0.5969, 0.5927])
    close4(nn.layer(3).numpy(), [1.4022, 1.4914])
    close4(nn.w(3).numpy(), [0.1585, 0.1500])
    close4(nn.layer(4).numpy(), [0.8025, 0.8163])
    tgt = th.Tensor.vector(2, [0.01, 0.99])
    close4(nn.loss(t4.LOSS_MSE, tgt), 0.658292, 2e-6)
    nn.backprop(tgt)
    close4(nn.layer(4).numpy(), [0.7925, -0.1737]); close4(nn.layer(3).numpy(), [0.7925, -0.1737])
    close4(nn.db(2).numpy(), [0.7925, -0.1737])
    close4(nn.dw(2).numpy(), [0.4702, 0.4731, 0.4697, -0.1031, -0.1037, -0.1029])
    close4(nn.layer(2).numpy(), [0.2215, 0.2698, 0.3181]); close4(nn.layer(1).numpy(), [0.2215, 0.2698, 0.3181])
    close4(nn.db(0).numpy(), [0.2215, 0.2698, 0.3181])
    close4(nn.dw(0).numpy(), [0.0111, 0.0221, 0.0135, 0.0270, 0.0159, 0.0318])
    close4(nn.layer(0).numpy(), [0.1643, 0.1729])
    nn.sgd(0.5, 0.0)
    close4(nn.w(2).numpy(), [0.1649, 0.2135, 0.2651, 0.6015, 0.5518, 0.5015])
    close4(nn.b(2).numpy(), [0.2037, 0.6869])
    assert not nn.dw(2).numpy().any() and not nn.db(2).numpy().any()
    close4(nn.w(0).numpy(), [0.1445, 0.1889, 0.2433, 0.2865, 0.1920, 0.1341])      # the script's `verify` line
    close4(nn.b(0).numpy(), [0.2393, 0.2151, 0.1909])


def test_t4_30c_mazur_n3():
    nn = mazur(3, 2)
    nn.set_w(0, [0.15, 0.2, 0.25, 0.3]).set_b(0, [0.35] * 2).set_w(2, [0.4, 0.45, 0.5, 0.55]).set_b(2, [0.6] * 2)
    nn.forward(th.Tensor.vector(6, [0.05, 0.1] * 3))                          # auto-reshaped: numel check only
    close4(nn.layer(4).numpy(), [0.7514, 0.7729] * 3)
    tgt = th.Tensor.vector(6, [0.01, 0.99] * 3).reshape(3, 1, 2, 1)
    close4(nn.loss(t4.LOSS_MSE, tgt), 0.596742, 2e-6)
    nn.backprop(tgt)
    close4(nn.db(0).numpy(), [0.5640, 0.6427]); close4(nn.dw(0).numpy(), [0.0282, 0.0564, 0.0321, 0.0643])
    close4(nn.layer(0).numpy(), [0.0818, 0.1019] * 3)
    nn.sgd(0.5, 0.0)
    close4(nn.w(0).numpy(), [0.1359, 0.1718, 0.2339, 0.2679]); close4(nn.b(0).numpy(), [0.0680, 0.0287])


# ------------------------------------------------------------------ model vs oracle with injected parameters
def inject(gm, om):
    """copy the oracle model's (seeded) parameters into the GPU model with nn.w= / nn.b="""
    for i, L in enumerate(om.layers[:-1]):
        if L.w is not None and L.dw is not None:
            gm.set_w(i, L.w.ravel()); gm.set_b(i, L.b.ravel())


def compare_layers(gm, om, what, rtol=1e-4):
    for i, L in enumerate(om.layers):
        assert_close(gm.layer(i).numpy(), L.data, rtol=rtol, what="%s layer %d" % (what, i))


def fragile_rows(om, rel=2e-5):
    """samples of the batch the oracle just ran forward whose result hangs on a rounding error: a relu / leaky-relu / elu / selu pre-activation
    within `rel` x rms of its kink, or a max/min-pool window whose two leading entries are that close (the routed index would flip).  FP32
    kernels with another summation order differ by ~1e-6 rel there, and a flipped mask is a DISCRETE change of the sample's whole backward
    pass (measured: one flip among 786k leaky-relu units moves dX of that sample by 15 % and every upstream dW past the 1e-4 bar)."""
    bad = np.zeros(om.layers[0].data.shape[0], bool)
    for t in om.layers[:-1]:
        d = t.data
        thr = rel * (np.sqrt(np.mean(d.astype(np.float64) ** 2)) + 1e-30)
        if t.fn in (orc.L_RELU, orc.L_LEAKYRL, orc.L_ELU, orc.L_SELU):
            bad |= (np.abs(d) < thr).reshape(d.shape[0], -1).any(axis=1)
        elif t.fn in (orc.L_MAXPOOL, orc.L_MINPOOL) and d.shape[1] % t.K == 0 and d.shape[2] % t.K == 0:
            n, h, w, c = d.shape
            win = d.reshape(n, h // t.K, t.K, w // t.K, t.K, c).transpose(0, 1, 3, 5, 2, 4).reshape(n, -1, t.K * t.K)
            srt = np.sort(win, axis=2)
            gap = (srt[:, :, -1] - srt[:, :, -2]) if t.fn == orc.L_MAXPOOL else (srt[:, :, 1] - srt[:, :, 0])
            bad |= (gap < thr).any(axis=1)
    return np.nonzero(bad)[0]


def settled_batch(om, draw, tries=30):
    """a batch none of whose samples sits on a kink (fragile_rows): the fragile samples are drawn again.  draw(k) -> k fresh samples."""
    x = draw(om.layers[0].data.shape[0])
    for _ in range(tries):
        om.forward(x)
        rows = fragile_rows(om)
        if rows.size == 0:
            return x
        x[rows] = draw(rows.size)
    raise AssertionError("no settled batch after %d tries" % tries)


def adam_slack(gm, om, lr, b1=0.9, b2=0.999, noise_rel=1e-4, first=True):
    """per-element slack on the parameters after an Adam step, from the gradients BEFORE it.  Adam without bias correction (nmath.cu:438-454)
    moves a parameter by u = lr*m/(sqrt(v)+1e-6) with m >= (1-b1)*g and v >= (1-b2)*g^2 in magnitude: du/dg <= lr*2*(1-b1)/(sqrt(1-b2)*|g|+1e-6), i.e. a
    sign-like update whose slope near g = 0 is lr*1e5 — FP32 summation-order noise dg in a gradient that happens to lie near zero (a handful of the
    401408 elements of a 784x512 layer at N=1024) shows as up to 2*lr*(1-b1)/sqrt(1-b2) in the parameter for ANY FP32 kernel.  dg is measured
    (max |dG - dG_oracle|, itself held to the 1e-4 bar by compare_params) or, when the step under test never exposes its gradients (gm None: the
    fused / captured steps), taken as the bar itself, noise_rel x rms(dG); everything away from g = 0 keeps the plain tolerance.  After the
    first step the moments carry the earlier steps' noise as well (only the parameters are re-synchronised between steps): m's noise is
    bounded by (1-b1) * sum b1^k * dg <= dg, hence the factor 1 instead of (1-b1) when not `first`."""
    c = (1.0 - b1) if first else 1.0
    out = {}
    for i, L in enumerate(om.layers[:-1]):
        if L.w is not None and L.dw is not None:
            sl = []
            for k, ref in enumerate((L.dw, L.db)):
                ref = np.asarray(ref, np.float64).ravel()
                if gm is not None:
                    got = gm.dw(i).numpy() if k == 0 else gm.db(i).numpy()
                    noise = float(np.abs(np.asarray(got, np.float64).ravel() - ref).max())
                else:
                    noise = noise_rel * float(np.sqrt(np.mean(ref * ref)))
                slope = lr * 2.0 * c / (np.sqrt(1.0 - b2) * np.maximum(np.abs(ref) - noise, 0.0) + 1e-6)
                sl.append(np.minimum(slope * noise, 2.0 * lr * (1.0 - b1) / np.sqrt(1.0 - b2)))
            out[i] = sl
    return out


def compare_params(gm, om, what, grads=True, rtol=1e-4, w_atol=0.0, slack=None, stragglers=0.0, straggler_cap=0.0):
    """w_atol: absolute slack on weights after an Adam step.  Adam without bias correction
    (nmath.cu:438-454) moves a weight by lr*m/(sqrt(v)+1e-6); for |dg| ~ 1e-5 that quotient has
    slope 1e5 in dg, so FP32 summation-order noise of 1e-8 in a gradient shows up as lr*1e-3 in w.
    Gradients whose exact value is 0 (conv bias in front of a batch-norm) are pure rounding noise:
    compared with an absolute floor tied to the layer's weight-gradient scale.
    slack: per-element version of the same argument (adam_slack).  stragglers: fraction of a tensor's elements that may exceed even that, as long
    as they stay within straggler_cap (the largest move one Adam step can make, 2*lr*(1-b1)/sqrt(1-b2)) — for steps that never expose their
    gradients, where the gradient noise entering adam_slack is an estimate (a few elements in 400k whose sums cancel harder than the layer's rms tells)."""
    def close(got, ref, atol, nm):
        if stragglers <= 0.0:
            return assert_close(got, ref, rtol=rtol, atol=atol, what=nm)
        g, r = np.asarray(got, np.float64).ravel(), np.asarray(ref, np.float64).ravel()
        err = np.abs(g - r)
        bad = err > rtol * np.abs(r) + atol
        assert bad.sum() <= max(1, int(stragglers * r.size)) and (err[bad] <= straggler_cap + rtol * np.abs(r[bad])).all(), \
            "%s: %d/%d out of tolerance, max abs err %.3e (cap %.3e)" % (nm, int(bad.sum()), r.size, err.max(), straggler_cap)
    for i, L in enumerate(om.layers[:-1]):
        if L.w is not None and L.dw is not None:
            rw = np.sqrt(np.mean(L.w.astype(np.float64) ** 2))
            sw, sb = slack[i] if slack else (0.0, 0.0)
            close(gm.w(i).numpy(), L.w, rtol * rw + w_atol + sw, "%s w%d" % (what, i))
            close(gm.b(i).numpy(), L.b, rtol * rw + w_atol + sb, "%s b%d" % (what, i))
            if grads:
                floor = 1e-5 * (np.sqrt(np.mean(L.dw.astype(np.float64) ** 2)) + 1e-3)
                assert_close(gm.dw(i).numpy(), L.dw, rtol=rtol, what="%s dw%d" % (what, i))
                rb = np.sqrt(np.mean(L.db.astype(np.float64) ** 2))
                assert_close(gm.db(i).numpy(), L.db, rtol=rtol, atol=rtol * rb + floor, what="%s db%d" % (what, i))


def build_pair(kind, N):
    if kind == "mnist":               # examples/t4_40a.4th:10-13
        gm = th.mnist_cnn(N)
        om = orc.OracleModel(N, 28, 28, 1, seed=5)
        om.add(orc.L_CONV, 10, 0.5, [3, 1, 1, 1]).add(orc.L_MAXPOOL, 2).add(orc.L_RELU).add(orc.L_FLATTEN)
        om.add(orc.L_LINEAR, 100, 1.0).add(orc.L_RELU).add(orc.L_LINEAR, 10, 1.0).add(orc.L_SOFTMAX)
        shape, E, lop = (N, 28, 28, 1), 10, t4.LOSS_CE
    elif kind == "toycnn":            # examples/t4_30d.4th:3-24 without dropout (RNG is not parity-comparable)
        gm = (th.Model(N, 16, 16, 1).conv2d(0.5, 2).maxpool(2).relu().conv2d(0.5, 2).maxpool(2).relu()
              .flatten().linear(16, 0.0).linear(4, 0.0).softmax())
        om = orc.OracleModel(N, 16, 16, 1, seed=6)
        om.add(orc.L_CONV, 2, 0.5, [3, 1, 1, 1]).add(orc.L_MAXPOOL, 2).add(orc.L_RELU)
        om.add(orc.L_CONV, 2, 0.5, [3, 1, 1, 1]).add(orc.L_MAXPOOL, 2).add(orc.L_RELU)
        om.add(orc.L_FLATTEN).add(orc.L_LINEAR, 16, 0.0).add(orc.L_LINEAR, 4, 0.0).add(orc.L_SOFTMAX)
        shape, E, lop = (N, 16, 16, 1), 4, t4.LOSS_CE
    elif kind == "gan_g":             # examples/t4_40b.4th:44-48
        gm = th.gan_generator(N)
        om = orc.OracleModel(N, 128, 1, 1, seed=7)
        om.add(orc.L_LINEAR, 256, 1.0).add(orc.L_LEAKYRL, 0, 0.2).add(orc.L_LINEAR, 512, 1.0).add(orc.L_LEAKYRL, 0, 0.2)
        om.add(orc.L_LINEAR, 784, 1.0).add(orc.L_TANH)
        shape, E, lop = (N, 128, 1, 1), 784, t4.LOSS_MSE
    elif kind == "dcgan":             # conv-transpose up-sampling stack (word `dconv2d`, netvm.cpp:315): 4x4 -> 8x8 -> 16x16
        gm = th.Model(N, 4, 4, 8).dconv2d(0.5, 6).relu().dconv2d(0.5, 1).tanh()
        om = orc.OracleModel(N, 4, 4, 8, seed=9)
        om.add(orc.L_DCONV, 6, 0.5).add(orc.L_RELU).add(orc.L_DCONV, 1, 0.5).add(orc.L_TANH)
        shape, E, lop = (N, 4, 4, 8), 256, t4.LOSS_MSE
    elif kind == "bn":                # conv + batchnorm block as in examples/t4_30e.4th:28-31
        gm = (th.Model(N, 8, 8, 3).conv2d(0.5, 6).batchnorm().relu().avgpool(2).flatten().linear(5).sigmoid())
        om = orc.OracleModel(N, 8, 8, 3, seed=8)
        om.add(orc.L_CONV, 6, 0.5, [3, 1, 1, 1]).add(orc.L_BATCHNM, 0, 0.1).add(orc.L_RELU).add(orc.L_AVGPOOL, 2)
        om.add(orc.L_FLATTEN).add(orc.L_LINEAR, 5, 1.0).add(orc.L_SIGMOID)
        shape, E, lop = (N, 8, 8, 3), 5, t4.LOSS_BCE
    inject(gm, om)
    return gm, om, shape, E, lop


@pytest.mark.parametrize("kind,N", [("mnist", 8), ("mnist", 64), ("mnist", 512), ("toycnn", 2), ("gan_g", 16), ("gan_g", 1024), ("bn", 4), ("dcgan", 6)])   # 512 / 1024: the BASELINE batch sizes
@pytest.mark.parametrize("opt", ["sgd", "adam", "adamw"])
def test_model_train_steps_vs_oracle(kind, N, opt):
    rng = np.random.default_rng(11)
    gm, om, shape, E, lop = build_pair(kind, N)
    assert len(gm) == len(om.layers)
    for step in range(3):
        draw = lambda k: (rng.random((k,) + tuple(shape[1:]), dtype=np.float32) * 2 - 1).astype(np.float32)
        x = settled_batch(om, draw) if N >= 512 else draw(N)
        if lop == t4.LOSS_MSE:
            y = (rng.random((N, E), dtype=np.float32) * 2 - 1).astype(np.float32)
        else:
            y = orc.onehot(rng.integers(0, E, N), E)
        X, Y = th.Tensor.from_numpy(x), th.Tensor.tensor(N, 1, E, 1, y)
        gm.forward(X); om.forward(x)
        compare_layers(gm, om, "%s fwd step %d" % (kind, step))
        assert_close(gm.loss(lop, Y), om.loss(lop, y), rtol=1e-4, atol=1e-6, what="loss")    # north star: <= 1e-4 loss deviation
        gm.backprop(Y); om.backprop(y)
        compare_layers(gm, om, "%s bwd step %d" % (kind, step))
        compare_params(gm, om, "%s bwd step %d" % (kind, step))
        if opt == "sgd":
            # gradients are batch SUMS (backprop.cu:76-109 does not divide by N): at the BASELINE batch sizes lr = 0.05 throws the weights
            # far out (generator, N=1024: activations up to 60, dW elements = sums of +-25000 that cancel to ~3 — bench_scripts/dbg_gan1024.py),
            # where 1e-6 of operand noise is 4e-3 in dW for ANY FP32 kernel, the reference's included; keep the step size per sample
            lr = 0.05 * min(1.0, 64.0 / N)
            gm.sgd(lr, 0.9); om.sgd(lr, 0.9)                  # momentum is forced to 0 on the first call (gradient.cu:139)
        elif opt == "adam":
            slack = adam_slack(gm, om, 0.001, first=step == 0)
            gm.adam(0.001); om.adam(0.001)
        else:
            slack = adam_slack(gm, om, 0.001, first=step == 0)
            gm.adamw(0.001, 0.01); om.adamw(0.001, 0.01)
        compare_params(gm, om, "%s %s step %d" % (kind, opt, step), grads=False, w_atol=0.0 if opt == "sgd" else 0.05 * 0.001,
                       slack=None if opt == "sgd" else slack)
        for i, L in enumerate(om.layers[:-1]):
            if L.dw is not None and L.w is not None:
                assert not gm.dw(i).numpy().any() and not gm.db(i).numpy().any()     # optimizers zero dG
                # next step starts from identical parameters (Adam's 1/(sqrt(v)+eps) amplifies rounding noise, see compare_params)
                L.w[...] = gm.w(i).numpy().reshape(L.w.shape); L.b[...] = gm.b(i).numpy().reshape(L.b.shape)


def test_trainable_off_and_hit():
    N = 16
    gm, om, shape, E, lop = build_pair("mnist", N)
    rng = np.random.default_rng(2)
    x = rng.random(shape, dtype=np.float32); lab = rng.integers(0, E, N).astype(np.int32)
    y = orc.onehot(lab, E)
    gm.trainable(False); om.train = False
    gm.forward(th.Tensor.from_numpy(x)); om.forward(x)
    dl = torch.from_numpy(lab).cuda()
    gm.onehot_labels(C.c_void_p(dl.data_ptr()))
    assert gm.hit() == orc.hit(om.output().reshape(N, E), y)
    assert_close(gm.loss(lop), om.loss(lop, y), rtol=1e-4)
    gm.backprop(); om.backprop(y)
    gm.adam(0.001); om.adam(0.001)
    compare_params(gm, om, "not trainable")                   # dW stays 0, weights unchanged
    compare_layers(gm, om, "not trainable bwd")


def test_errors_mirror_reference():
    nn = th.mnist_cnn(4)
    with pytest.raises(t4.T4KError, match="wrong shape"):
        nn.forward(th.Tensor.tensor(4, 28, 28, 2))            # forward.cu:33-38
    with pytest.raises(t4.T4KError, match="Onehot wrong shape"):
        nn.backprop(th.Tensor.vector(7))                      # backprop.cu:78-83
    with pytest.raises(t4.T4KError):
        th.Model(2, 8, 8, 1).maxpool(4)                       # model.cpp:262-265
    with pytest.raises(t4.T4KError):
        th.Model(2, 8, 8, 1).conv2d(0.5, 4, k=7)              # model.cpp:144-149


def test_step_graph_equals_eager():
    N = 32
    rng = np.random.default_rng(4)
    x = (rng.random((N, 28, 28, 1), dtype=np.float32) * 2 - 1).astype(np.float32); y = orc.onehot(rng.integers(0, 10, N), 10)
    ga, om, *_ = build_pair("mnist", N)
    gb, _, *_ = build_pair("mnist", N)
    X, Y = th.Tensor.from_numpy(x), th.Tensor.tensor(N, 1, 10, 1, y)
    la, lb = torch.zeros(1, device="cuda"), torch.zeros(1, device="cuda")
    losses_a, losses_b = [], []
    for step in range(5):
        ga.forward(X); ga.loss_async(t4.LOSS_CE, Y, C.c_void_p(la.data_ptr())); ga.backprop(Y); ga.adam(0.001)
        th.sync(); losses_a.append(float(la.cpu()[0]))
        assert gb.step_graph(X, Y, t4.LOSS_CE, C.c_void_p(lb.data_ptr()), optimizer=2, lr=0.001) == 0
        th.sync(); losses_b.append(float(lb.cpu()[0]))
    # the captured step takes the train tail (head backward inside the forward tail kernel, head parameter gradients summed from per-CTA
    # partials) and the one-launch dX/dW pair: same arithmetic, other summation orders -> equal within FP32 rounding noise
    assert np.allclose(losses_a, losses_b, rtol=2e-6, atol=0), (losses_a, losses_b)
    for i in (0, 4, 6):
        assert_close(gb.w(i).numpy(), ga.w(i).numpy(), rtol=2e-5, what="weights of layer %d" % i)


@pytest.mark.parametrize("N", [64, 512])                     # 512: BASELINE config 3, the shape bench.py times
def test_step_graph_trajectory_vs_oracle(N):
    """the CAPTURED train step (train tail, one-launch dX/dW pair, optimizer split — the path bench.py measures) against the oracle's Model
    restatement: loss trajectory within the north star's 1e-4 and parameters after every Adam step"""
    rng = np.random.default_rng(17)
    gm, om, shape, E, lop = build_pair("mnist", N)
    ld = torch.zeros(1, device="cuda")
    lp = C.c_void_p(ld.data_ptr())
    x = (rng.random(shape, dtype=np.float32) * 2 - 1).astype(np.float32); y = orc.onehot(rng.integers(0, E, N), E)
    X, Y = th.Tensor.from_numpy(x), th.Tensor.tensor(N, 1, E, 1, y)
    for step in range(6):                                      # steps 0-1 run eagerly (arenas, first optimizer call), then capture + replays
        assert gm.step_graph(X, Y, lop, lp, optimizer=2, lr=0.001) == 0
        th.sync()
        om.forward(x); lo = om.loss(lop, y); om.backprop(y)
        slack = adam_slack(None, om, 0.001, first=step == 0)
        om.adam(0.001)
        assert_close(float(ld.cpu()[0]), lo, rtol=1e-4, atol=1e-6, what="loss of step %d" % step)
        compare_params(gm, om, "step_graph step %d" % step, grads=False, w_atol=0.05 * 0.001, slack=slack, stragglers=1e-4, straggler_cap=2 * 0.001 * 0.1 / np.sqrt(1e-3))
        for i, L in enumerate(om.layers[:-1]):
            if L.dw is not None and L.w is not None:
                assert not gm.dw(i).numpy().any() and not gm.db(i).numpy().any()
                L.w[...] = gm.w(i).numpy().reshape(L.w.shape); L.b[...] = gm.b(i).numpy().reshape(L.b.shape)


@pytest.mark.parametrize("kind,N", [("mnist", 32), ("toycnn", 3)])
def test_fused_block_equals_per_layer(kind, N):
    """conv → maxpool(2) → relu (→ flatten) fused kernels write the same layer tensors as the per-layer launches:
    bit-equal activations / routed gradients / dX (same arithmetic order), dF/dB within summation-order noise"""
    rng = np.random.default_rng(21)
    ga, om, shape, E, lop = build_pair(kind, N)
    gb, _, *_ = build_pair(kind, N)
    gb.fuse(False)
    x = (rng.random(shape, dtype=np.float32) * 2 - 1).astype(np.float32); y = orc.onehot(rng.integers(0, E, N), E)
    X, Y = th.Tensor.from_numpy(x), th.Tensor.tensor(N, 1, E, 1, y)
    ga.forward(X); gb.forward(X)
    first_fc = min(i for i, t in enumerate(om.layers) if t.fn == orc.L_LINEAR)
    for i in range(len(ga)):
        a, b = ga.layer(i).numpy(), gb.layer(i).numpy()
        if i > first_fc:
            assert_close(a, b, rtol=1e-5, what="fwd layer %d" % i)          # FC tail: split-K / dot-product summation order differs
        else:
            assert np.array_equal(a, b), "fwd layer %d" % i                  # conv blocks: same FMA order, pool routing, relu: bit-equal
    n0 = t4.load().t4k_launch_count()
    ga.forward(X); ga.backprop(Y)
    n1 = t4.load().t4k_launch_count()
    gb.forward(X); gb.backprop(Y)
    n2 = t4.load().t4k_launch_count()
    assert n1 - n0 < n2 - n1                                   # fewer launches
    for i in range(len(ga)):
        a, b = ga.layer(i).numpy(), gb.layer(i).numpy()
        assert_close(a, b, rtol=2e-5, what="bwd layer %d" % i)
        if i == 1:
            assert np.array_equal(a == 0, b == 0), "arg-max routing pattern of the conv output gradient"
    for i in range(len(ga) - 1):
        if ga.dw(i) is not None and ga.w(i) is not None and ga.db(i) is not None:
            assert_close(ga.dw(i).numpy(), gb.dw(i).numpy(), rtol=1e-5, what="dw%d" % i)
            assert_close(ga.db(i).numpy(), gb.db(i).numpy(), rtol=1e-5, what="db%d" % i)


# ------------------------------------------------------------------ Dataset feeding (src/mu/dataset.cu, forward.cu:72-75)
def test_dataset_feed_forward_onehot_hit_and_train_step():
    """U8 mini-batches staged asynchronously (double buffered), normalised on the device, fed to Model::forward(Dataset&):
    same tensor, one-hot, hit count and training trajectory as float batches prepared on the host the way Dataset::_load does."""
    N, E = 32, 10
    rng = np.random.default_rng(3)
    batches = [(rng.integers(0, 256, (N, 28, 28, 1), dtype=np.uint8), rng.integers(0, 10, N, dtype=np.uint8)) for _ in range(4)]
    ds = th.Dataset(N, 28, 28, 1).normalize(128.0, 128.0)                 # t4_40b.4th:52 `128 128 normalize`
    mean, scale = orc.dataset_normalize(128.0, 128.0)

    def build():
        t4.load().t4k_rand_seed(77)
        return th.mnist_cnn(N)
    m, ref = build(), build()
    loss_dev = torch.zeros(8, device="cuda"); lp = C.c_void_p(loss_dev.data_ptr())
    loss_ref = torch.zeros(8, device="cuda"); lr_ = C.c_void_p(loss_ref.data_ptr())
    # forward(Dataset): tensor contents, one-hot and hit
    ds.stage(*batches[0]).commit()
    x0 = orc.dataset_load(batches[0][0], mean, scale).reshape(N, 28, 28, 1)
    assert np.array_equal(ds.tensor.numpy().view(np.uint32).ravel(), x0.view(np.uint32).ravel())
    m.forward_ds(ds)
    hot0 = orc.onehot(batches[0][1].astype(np.int32), E)
    assert m.hit(False) == orc.hit(m.layer(-1).numpy().reshape(N, E), hot0)
    ref.forward(th.Tensor.from_numpy(x0))
    assert np.array_equal(m.layer(-1).numpy(), ref.layer(-1).numpy())
    # training: stage batch i+1 while batch i trains
    ds.stage(*batches[0])
    got, want = [], []
    for i in range(4):
        if i + 1 < 4:
            ds.stage(*batches[i + 1])
        t4.check(m.step_graph_ds(ds, t4.LOSS_CE, lp, optimizer=2, lr=1e-3), "step_graph_ds")
        th.sync(); got.append(float(loss_dev[0].cpu()))
        # the load rides in the step's first fused block: the dataset tensor still holds exactly what Dataset::_load would have put there
        assert np.array_equal(ds.tensor.numpy().view(np.uint32).ravel(), orc.dataset_load(batches[i][0], mean, scale).view(np.uint32).ravel())
        xi = th.Tensor.from_numpy(orc.dataset_load(batches[i][0], mean, scale).reshape(N, 28, 28, 1))
        yi = th.Tensor.tensor(N, 1, E, 1, orc.onehot(batches[i][1].astype(np.int32), E))
        ref.forward(xi); ref.loss_async(t4.LOSS_CE, yi, lr_); ref.backprop(yi); ref.adam(1e-3)
        th.sync(); want.append(float(loss_ref[0].cpu()))
    np.testing.assert_allclose(got, want, rtol=2e-6)
    # the captured step takes the train tail and the one-launch dX/dW pair (other summation orders than the eager per-layer calls): FP32 noise,
    # which Adam's sign-like update turns into a visible move only where a gradient lies within rounding of 0 (see adam_slack)
    dw = np.abs(m.w(4).numpy().astype(np.float64) - ref.w(4).numpy()).ravel()
    assert np.mean(dw > 1e-6) < 1e-3 and dw.max() <= 4 * 2e-3 * 3.2, (np.mean(dw > 1e-6), dw.max())
    # train_step_ds: same step, loss read-back pipelined by one call
    ds.stage(*batches[0])
    seen = []
    for i in range(3):
        if i + 1 < 3:
            ds.stage(*batches[i + 1])
        seen.append(m.train_step_ds(ds, t4.LOSS_CE, lp, optimizer=2, lr=1e-3))
        xi = th.Tensor.from_numpy(orc.dataset_load(batches[i][0], mean, scale).reshape(N, 28, 28, 1))
        yi = th.Tensor.tensor(N, 1, E, 1, orc.onehot(batches[i][1].astype(np.int32), E))
        ref.forward(xi); ref.loss_async(t4.LOSS_CE, yi, lr_); ref.backprop(yi); ref.adam(1e-3)
        th.sync(); want.append(float(loss_ref[0].cpu()))
    seen.append(m.train_flush())
    assert np.isnan(seen[0])
    np.testing.assert_allclose(seen[1:], want[4:], rtol=1e-6)


# ------------------------------------------------------------------ GAN iteration (BASELINE config 4; examples/t4_40b.4th:37-67)
@pytest.mark.parametrize("N", [16, 1024])                   # 1024: BASELINE config 4
def test_gan_iteration_vs_oracle(N):
    """train_d + train_g exactly as the script sequences them (two accumulated D backprops -> Adam(b1=0.5); D frozen, dX of D's
    input back-propagated through G -> Adam) against the oracle's Model restatement.  Dropout p = 0 here: the mask RNG is not
    parity-comparable; the layer itself stays in the path (mask = all ones)."""
    rng = np.random.default_rng(21)
    D, G = th.gan_discriminator(N, p=0.0), th.gan_generator(N)
    oD = orc.OracleModel(N, 28, 28, 1, seed=31)
    oD.add(orc.L_LINEAR, 512, 1.0).add(orc.L_LEAKYRL, 0, 0.2).add(orc.L_DROPOUT, 0, 0.0).add(orc.L_LINEAR, 256, 1.0).add(orc.L_LEAKYRL, 0, 0.2)
    oD.add(orc.L_DROPOUT, 0, 0.0).add(orc.L_LINEAR, 1, 1.0).add(orc.L_SIGMOID)
    oG = orc.OracleModel(N, 128, 1, 1, seed=32)
    oG.add(orc.L_LINEAR, 256, 1.0).add(orc.L_LEAKYRL, 0, 0.2).add(orc.L_LINEAR, 512, 1.0).add(orc.L_LEAKYRL, 0, 0.2).add(orc.L_LINEAR, 784, 1.0).add(orc.L_TANH)
    inject(D, oD); inject(G, oG)
    ones, zeros_ = np.ones((N, 1), np.float32), np.zeros((N, 1), np.float32)
    REAL, FAKE = th.Tensor.tensor(N, 1, 1, 1, ones), th.Tensor.tensor(N, 1, 1, 1, zeros_)
    for it in range(2):
        real = (rng.random((N, 28, 28, 1), dtype=np.float32) * 2 - 1).astype(np.float32)
        z1, z2 = rng.standard_normal((N, 128, 1, 1)).astype(np.float32), rng.standard_normal((N, 128, 1, 1)).astype(np.float32)
        got = th.gan_iteration(D, G, th.Tensor.from_numpy(real), th.Tensor.from_numpy(z1), th.Tensor.from_numpy(z2), REAL, FAKE)
        # the same sequence on the oracle (t4_40b.4th:60-67)
        oD.train = True
        oD.forward(real); l_dr = oD.loss(orc.LOSS_BCE, ones); oD.backprop(ones)
        fake = oG.forward(z1).output().reshape(N, 28, 28, 1).copy()
        oD.forward(fake); l_df = oD.loss(orc.LOSS_BCE, zeros_); oD.backprop(zeros_)
        slD = adam_slack(None, oD, 1e-4, 0.5, first=it == 0)
        oD.adam(1e-4, 0.5)
        oD.train = False
        fake = oG.forward(z2).output().reshape(N, 28, 28, 1).copy()
        oD.forward(fake); l_gr = oD.loss(orc.LOSS_BCE, ones); oD.backprop(ones)
        oG.backprop(oD.layers[0].data.reshape(N, -1).copy())
        slG = adam_slack(None, oG, 4e-4, 0.5, first=it == 0)
        oG.adam(4e-4, 0.5)
        assert_close(got, (l_dr, l_df, l_gr), rtol=1e-4, atol=1e-6, what="GAN losses it %d" % it)
        cap = lambda lr: 2 * lr * 0.5 / np.sqrt(1e-3)
        compare_params(D, oD, "D it %d" % it, grads=False, w_atol=0.05 * 1e-4, slack=slD, stragglers=1e-4, straggler_cap=cap(1e-4))
        compare_params(G, oG, "G it %d" % it, grads=False, w_atol=0.05 * 4e-4, slack=slG, stragglers=1e-4, straggler_cap=cap(4e-4))
        for m_, o_ in ((D, oD), (G, oG)):                       # next iteration from identical parameters (Adam amplifies rounding noise)
            for i, L in enumerate(o_.layers[:-1]):
                if L.w is not None and L.dw is not None:
                    L.w[...] = m_.w(i).numpy().reshape(L.w.shape); L.b[...] = m_.b(i).numpy().reshape(L.b.shape)


# ------------------------------------------------------------------ model file (src/io/aio_model.cpp)
def _lit(v):
    """a number as the reference VM holds it: FP32 with bit 0 cleared (bit 0 tags objects, src/t4base.h:16-30 SCALAR)"""
    return (np.array([v], np.float32).view(np.uint32) & np.uint32(0xFFFFFFFE)).view(np.float32)[0]


def _chaos(n, scale):
    """integration/scripts/*.4th `chaos` in numpy: x0 = 0.1 + 0.8 j/n, 20 x (x <- 4 x (1 - x)), - 0.5, * scale — every step one IEEE
    FP32 operation (and every literal the VM's bit-0-cleared float), so the reference's tensor words produce the same bits"""
    f = np.float32
    j = np.arange(n, dtype=np.int64)
    x = (f(1.0) * j.astype(np.float32) / f(n)).astype(np.float32)          # gradfill: v * j / n (t4math.cu:192)
    x = (x * _lit(0.8)).astype(np.float32); x = (x + _lit(0.1)).astype(np.float32)
    for _ in range(20):
        y = ((x * f(-1.0)).astype(np.float32) + f(1.0)).astype(np.float32)
        x = ((x * y).astype(np.float32) * f(4.0)).astype(np.float32)
    return ((x - f(0.5)).astype(np.float32) * _lit(scale)).astype(np.float32)


def test_model_file_matches_the_reference_byte_for_byte(tmp_path):
    """Model::save writes the reference's model file: compared byte for byte with what the reference binary itself saves for the same
    model and weights (when oracle/_ref/ten4 is present), and reloaded into a fresh model (parameter path of AIO::nload)."""
    import os, subprocess
    N = 4
    def build():
        return th.Model(N, 8, 8, 1).conv2d(0.5, 2).maxpool(2).relu().flatten().linear(6, 0.0).batchnorm().leakyrelu(0.1).linear(3, 0.0).softmax()
    m = build()
    # layers: 0 conv2d, 1 maxpool, 2 relu, 3 flatten, 4 linear, 5 batchnm, 6 leakyrl, 7 linear, 8 softmax
    w = {(0, "w"): _chaos(18, 1.6), (0, "b"): _chaos(2, 0.2), (4, "w"): _chaos(6 * 32, 0.4), (4, "b"): _chaos(6, 0.2),
         (7, "w"): _chaos(18, 0.5), (7, "b"): _chaos(3, 0.2)}
    for (i, k), v in w.items():
        (m.set_w if k == "w" else m.set_b)(i, v)
    mine = str(tmp_path / "mine.t4")
    m.save(mine)
    data = open(mine, "rb").read()
    assert data.startswith(b"\\ tensorForth v4.0 model\nbias=0.5, C=2, K=3, S=1, P=1conv2d \n2x2maxpool\nrelu   \nflatten\nbias=0, H=6linear \n")
    assert data.endswith(b"\n---\n") and b"\n--- w.conv2d \n" + w[(0, "w")].tobytes() + b"\n--- b.conv2d \n" in data
    # round trip into a fresh model
    m2 = build().load(mine)
    for (i, k) in w:
        assert np.array_equal((m2.w(i) if k == "w" else m2.b(i)).numpy().ravel(), w[(i, k)])
    assert np.array_equal(m2.w(5).numpy(), m.w(5).numpy())                  # batchnorm gamma
    ref_bin = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "ten4")
    if not os.path.exists(ref_bin):
        return
    theirs = str(tmp_path / "ref.t4")
    script = "\n".join([
        "0 trace", ": lg copy -1 *= 1 += *= 4 *= ;", ": chaos gradfill 0.8 *= 0.1 += 19 for lg next 0.5 -= ;",
        "%d 8 8 1 nn.model 0.5 2 conv2d 2 maxpool relu flatten 0.0 6 linear 0.1 batchnorm 0.1 leakyrelu 0.0 3 linear softmax constant md" % N,
        "md", "1 3 3 2 tensor chaos 1.6 *= 0 nn.w=", "2 vector chaos 0.2 *= 0 nn.b=", "6 32 matrix chaos 0.4 *= 4 nn.w=", "6 vector chaos 0.2 *= 4 nn.b=",
        "3 6 matrix chaos 0.5 *= 7 nn.w=", "3 vector chaos 0.2 *= 7 nn.b=", 's" %s" save' % theirs, "drop", "bye", ""])
    p = subprocess.run([ref_bin], input=script, capture_output=True, text=True, timeout=120)
    assert os.path.exists(theirs), p.stdout[-1500:] + p.stderr[-500:]
    ref = open(theirs, "rb").read()
    assert ref == data, "model file differs from the reference's: %r ... vs %r ..." % (ref[:200], data[:200])


def test_step_graph_captured_on_its_first_call():
    """bench.py's order: eager steps first (arenas built, _iter > 0), so the FIRST step_graph call captures — nothing the step needs
    (workspaces, the softmax duplicate for the side-stream loss) may be allocated inside the capture"""
    N = 16
    rng = np.random.default_rng(9)
    x = (rng.random((N, 28, 28, 1), dtype=np.float32) * 2 - 1).astype(np.float32); y = orc.onehot(rng.integers(0, 10, N), 10)
    ga, _, *_ = build_pair("mnist", N)
    gb, _, *_ = build_pair("mnist", N)
    X, Y = th.Tensor.from_numpy(x), th.Tensor.tensor(N, 1, 10, 1, y)
    la, lb = torch.zeros(1, device="cuda"), torch.zeros(1, device="cuda")
    pa, pb = C.c_void_p(la.data_ptr()), C.c_void_p(lb.data_ptr())
    for m_, p_ in ((ga, pa), (gb, pb)):
        for _ in range(2):
            m_.forward(X); m_.loss_async(t4.LOSS_CE, Y, p_); m_.backprop(Y); m_.adam(0.001)
    for step in range(4):
        ga.forward(X); ga.loss_async(t4.LOSS_CE, Y, pa); ga.backprop(Y); ga.adam(0.001)
        assert gb.step_graph(X, Y, t4.LOSS_CE, pb, optimizer=2, lr=0.001) == 0
        th.sync()
        assert float(la.cpu()[0]) == float(lb.cpu()[0])
    assert np.array_equal(ga.w(4).numpy(), gb.w(4).numpy())


def test_graph_replays_draw_fresh_random_numbers():
    """th.Graph: a captured sequence that draws (randn, dropout masks) starts with the RNG replay-epoch tick, so every replay gets new
    numbers, and the captured computation equals the eager one on the same inputs"""
    t4.load().t4k_rand_seed(99)
    z = th.Tensor.tensor(8, 16, 1, 1)
    m = th.Model(8, 16, 1, 1).linear(12).leakyrelu(0.2).dropout(0.5).linear(4).sigmoid()
    x = th.Tensor.from_numpy(np.random.default_rng(2).standard_normal((8, 16, 1, 1)).astype(np.float32))
    m.forward(x)                                     # warm-up: workspaces exist before the capture

    def seq():
        z.randn(); m.forward(z)
    seq()
    g = th.Graph(seq)
    draws, masks = [], []
    for _ in range(3):
        g(); th.sync()
        draws.append(z.numpy().copy()); masks.append(m.ex(2).numpy().copy())     # layer 2 = dropout: its saved mask
    assert not np.array_equal(draws[0], draws[1]) and not np.array_equal(draws[1], draws[2])
    assert not np.array_equal(masks[0], masks[1])
    assert abs(float(np.mean(draws[2]))) < 0.5 and 0.5 < float(np.std(draws[2])) < 1.5
    # the graph computes what the eager path computes for the input it drew (dropout mask taken from the replay)
    out_g = m.layer(-1).numpy().copy()
    h = np.maximum(draws[2].reshape(8, 16) @ m.w(0).numpy().reshape(12, 16).T + m.b(0).numpy(), 0.2 * (draws[2].reshape(8, 16) @ m.w(0).numpy().reshape(12, 16).T + m.b(0).numpy()))
    h = h * masks[2].reshape(8, 12)
    o = 1.0 / (1.0 + np.exp(-(h @ m.w(3).numpy().reshape(4, 12).T + m.b(3).numpy())))
    np.testing.assert_allclose(out_g.reshape(8, 4), o, rtol=1e-4, atol=1e-6)


def test_step_graph_with_dropout_draws_a_new_mask_every_step():
    N = 16
    t4.load().t4k_rand_seed(5)
    m = th.Model(N, 8, 8, 1).flatten().linear(32).relu().dropout(0.5).linear(4).softmax()
    rng = np.random.default_rng(1)
    X = th.Tensor.from_numpy((rng.random((N, 8, 8, 1), dtype=np.float32) * 2 - 1).astype(np.float32))
    Y = th.Tensor.tensor(N, 1, 4, 1, orc.onehot(rng.integers(0, 4, N), 4))
    ld = torch.zeros(1, device="cuda"); lp = C.c_void_p(ld.data_ptr())
    masks, losses = [], []
    for step in range(5):
        assert m.step_graph(X, Y, t4.LOSS_CE, lp, optimizer=2, lr=1e-3) == 0
        th.sync(); masks.append(m.ex(3).numpy().copy()); losses.append(float(ld.cpu()[0]))
    assert all(np.isfinite(losses))
    assert set(np.unique(masks[-1])) <= {0.0, 1.0} and 0.2 < masks[-1].mean() < 0.8
    assert not np.array_equal(masks[2], masks[3]) and not np.array_equal(masks[3], masks[4])      # steps 3.. are graph replays


# ------------------------------------------------------------------ golden: what the reference PROGRAM printed (tests/golden/ref_program)
def _golden_numbers(name, key):
    import os, re
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_program", name + ".out")
    return [float(m) for m in re.findall(r"^%s=([-+0-9.eE]+)" % re.escape(key), open(path).read(), flags=re.M)]


def _close_printed(got, want, what):
    """the VM prints 6 significant digits; the north star's bar is 1e-4 relative"""
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    assert got.shape == want.shape, (what, got.shape, want.shape)
    assert np.all(np.abs(got - want) <= 1e-4 * np.abs(want) + 1e-5), (what, got.tolist(), want.tolist())


def test_reference_program_golden_cnn_parity():
    """integration/scripts/cnn_parity.4th through the host mirror: same model (t4_40a.4th:10-13 at N=64), same bits in weights, input and
    labels; loss before training, after one Adam step at lr=1e-3 and nine at 2e-5 — against what the reference program printed"""
    N = 64
    m = (th.Model(N, 28, 28, 1).conv2d(0.5, 10).maxpool(2).relu().flatten().linear(100, 0.0).relu().linear(10, 0.0).softmax())
    m.set_w(0, _chaos(90, 1.6)); m.set_b(0, _chaos(10, 0.2))
    m.set_w(4, _chaos(100 * 1960, 0.05)); m.set_b(4, _chaos(100, 0.1))
    m.set_w(6, _chaos(1000, 0.2)); m.set_b(6, _chaos(10, 0.1))
    x = (_chaos(N * 784, 1.0) * np.float32(2.0)).astype(np.float32).reshape(N, 28, 28, 1)
    y = np.zeros((N, 10), np.float32)
    y[np.arange(N), (7 * np.arange(N) + 3) % 10] = 1.0
    X, Y = th.Tensor.from_numpy(x), th.Tensor.tensor(N, 1, 10, 1, y)
    assert abs(float(x.astype(np.float64).sum()) - _golden_numbers("cnn_parity", "X sum")[0]) < 0.06        # printed with 4 digits: the input bits are the reference's
    m.forward(X)
    _close_printed([m.loss(t4.LOSS_CE, Y)], _golden_numbers("cnn_parity", "loss0"), "loss0")
    m.backprop(Y)
    _close_printed([np.sqrt(float((m.dw(6).numpy().astype(np.float64) ** 2).sum()))], _golden_numbers("cnn_parity", "dw6 norm"), "dw6 norm")
    m.adam(float(_lit(0.001)))
    losses = []
    for _ in range(10):
        m.forward(X); m.backprop(Y); m.adam(float(_lit(0.00002)))
        m.forward(X); losses.append(m.loss(t4.LOSS_CE, Y))
    _close_printed(losses, _golden_numbers("cnn_parity", "loss"), "loss trajectory")
    _close_printed([np.sqrt(float((m.w(6).numpy().astype(np.float64) ** 2).sum()))], _golden_numbers("cnn_parity", "w6 norm"), "w6 norm")


def test_reference_program_golden_mlp_bn():
    """integration/scripts/mlp_bn_parity.4th: flatten, linear, batchnorm, tanh, leakyrelu, sigmoid, MSE, SGD with momentum"""
    N = 32
    m = (th.Model(N, 8, 8, 2).flatten().linear(48, 0.0).batchnorm(0.1).tanh().linear(24, 0.0).leakyrelu(float(_lit(0.1))).linear(4, 0.0).sigmoid())
    m.set_w(1, _chaos(48 * 128, 0.3)); m.set_b(1, _chaos(48, 0.2))
    m.set_w(4, _chaos(24 * 48, 0.4)); m.set_b(4, _chaos(24, 0.2))
    m.set_w(6, _chaos(4 * 24, 0.5)); m.set_b(6, _chaos(4, 0.2))
    x = (_chaos(N * 128, 1.0) * np.float32(2.0)).astype(np.float32).reshape(N, 8, 8, 2)
    t = (_chaos(N * 4, 1.0) + _lit(0.5)).astype(np.float32).reshape(N, 4)
    X, T = th.Tensor.from_numpy(x), th.Tensor.tensor(N, 1, 4, 1, t)
    m.forward(X)
    _close_printed([m.loss(t4.LOSS_MSE, T)], _golden_numbers("mlp_bn_parity", "loss0"), "loss0")
    m.backprop(T)
    _close_printed([np.sqrt(float((m.dw(1).numpy().astype(np.float64) ** 2).sum()))], _golden_numbers("mlp_bn_parity", "dw1 norm"), "dw1 norm")
    m.sgd(float(_lit(0.01)), float(_lit(0.9)))
    losses = []
    for _ in range(10):
        m.forward(X); m.backprop(T); m.sgd(float(_lit(0.01)), float(_lit(0.9)))
        m.forward(X); losses.append(m.loss(t4.LOSS_MSE, T))
    _close_printed(losses, _golden_numbers("mlp_bn_parity", "loss"), "loss trajectory")
    _close_printed([np.sqrt(float((m.w(1).numpy().astype(np.float64) ** 2).sum()))], _golden_numbers("mlp_bn_parity", "w1 norm"), "w1 norm")


def test_model_file_carries_optimizer_state_for_resume(tmp_path):
    """save(opt_state=True) appends Adam's moment arenas and the step count behind the reference's sections; a fresh model that
    loads the file continues the run bit for bit, one that loads the parameters only (the reference's file) restarts from m = v = 0 and drifts;
    the parameter part of both files is the same bytes."""
    N = 16
    rng = np.random.default_rng(12)
    xs = [(rng.random((N, 28, 28, 1), dtype=np.float32) * 2 - 1).astype(np.float32) for _ in range(5)]
    ys = [orc.onehot(rng.integers(0, 10, N), 10) for _ in range(5)]

    def step(m, k):
        m.forward(th.Tensor.from_numpy(xs[k])); m.backprop(th.Tensor.tensor(N, 1, 10, 1, ys[k])); m.adam(1e-3)
    t4.load().t4k_rand_seed(5)
    a = th.mnist_cnn(N)
    for k in range(3):
        step(a, k)
    th.sync()
    plain, full = str(tmp_path / "plain.t4"), str(tmp_path / "state.t4")
    a.save(plain); a.save(full, opt_state=True)
    dp, df = open(plain, "rb").read(), open(full, "rb").read()
    assert df.startswith(dp) and b"\n\\ optimizer state iter=3 epoch=0 floats=" in df[len(dp) - 1:] and df.endswith(b"\n---\n")
    b, c = th.mnist_cnn(N).load(full), th.mnist_cnn(N).load(plain)
    for k in (3, 4):
        step(a, k); step(b, k); step(c, k)
    th.sync()
    for i in (0, 4, 6):
        assert np.array_equal(b.w(i).numpy(), a.w(i).numpy()) and np.array_equal(b.b(i).numpy(), a.b(i).numpy()), "resumed run diverged at layer %d" % i
    assert not np.array_equal(c.w(4).numpy(), a.w(4).numpy())              # without the moments the continuation is another trajectory
    # a state that does not fit the model is refused
    with pytest.raises(Exception):
        th.Model(N, 28, 28, 1).flatten().linear(10).softmax().load(full)
