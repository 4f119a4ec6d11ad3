"""
Pins the CPU oracle (oracle/) against every known-answer value the reference's own
example scripts hold for the hot path:
  examples/t4_20a.4th:2-9,45-68   matmul / += / -= / @= / Hadamard
  examples/t4_30a.4th:3-31        linear forward  -> {6,13,20}   (also README.md:326-331)
  examples/t4_30b.4th:3-69        Mazur 2-3-2 MLP, N=1: every activation, loss, dW/dB/dX, SGD
  examples/t4_30c.4th:4-70        same with N=3 (batch sums)
Printed precision of the reference is %+.4f, so comparisons use atol=5e-5+ (half a ulp
of the 4th decimal) unless the script gives more digits (loss: 6 decimals).
"""
import numpy as np
from oracle import oracle as orc

A4 = 6e-5          # half-ulp of the reference's 4-decimal print + float noise


def close(a, b, atol=A4):
    np.testing.assert_allclose(np.asarray(a, np.float64).ravel(), np.asarray(b, np.float64).ravel(), rtol=0, atol=atol)


# ------------------------------------------------------------------ t4_20a.4th
def test_t4_20a_matmul():
    A = np.array([[1, 2, 3], [4, 5, 6]], np.float32)
    B = np.ones((3, 2), np.float32)
    close(orc.gemm(A, B), [[6, 6], [15, 15]], 0)                       # :2-9


def test_t4_20a_add_sub():
    A = np.array([[1, 2, 3], [4, 5, 6]], np.float32)
    B = np.ones((2, 3), np.float32)
    S = orc.tt_op(orc.ADD, A, B)
    close(S, [[2, 3, 4], [5, 6, 7]], 0)                                # :45-51
    close(orc.tt_op(orc.SUB, S, np.full((2, 3), 2, np.float32)), [[0, 1, 2], [3, 4, 5]], 0)


def test_t4_20a_matmul_inplace_and_hadamard():
    A = np.array([[1, 2, 3], [0, 4, 5]], np.float32)
    C = orc.gemm(A, np.ones((3, 2), np.float32))
    close(C, [[6, 6], [9, 9]], 0)                                      # :57-62
    half = orc.ts_op(orc.MUL, np.ones((2, 2), np.float32), 0.5)
    close(orc.tt_op(orc.MUL, C, half), [[3, 3], [4.5, 4.5]], 0)        # :64-68


def test_t4_20a_large_ones():
    # :12-16  rand(512x1024) @ ones(1024x256) / 1024 : every column equals the row mean
    rng = np.random.default_rng(1)
    A = rng.random((64, 1024), dtype=np.float32)
    O = orc.ts_op(orc.DIV, orc.gemm(A, np.ones((1024, 256), np.float32)), 1024.0)
    close(O, np.repeat(A.astype(np.float64).mean(1)[:, None], 256, 1), 2e-6)


# ------------------------------------------------------------------ t4_30a.4th
def test_t4_30a_linear_forward():
    m = orc.OracleModel(1, 1, 2, 1).add(orc.L_LINEAR, 3, 1.0)
    m.layers[0].w[:] = 0.1 * np.array([[1, 2], [3, 4], [5, 6]], np.float32)
    m.layers[0].b[:] = [1, 2, 3]
    m.forward(np.array([10, 20], np.float32))
    close(m.output(), [6, 13, 20], 1e-5)


# ------------------------------------------------------------------ t4_30b.4th
def mazur(N, hidden):
    m = orc.OracleModel(N, 1, 2, 1)
    m.add(orc.L_LINEAR, hidden, 1.0).add(orc.L_SIGMOID)
    m.add(orc.L_LINEAR, 2, 1.0).add(orc.L_SIGMOID)
    return m


def test_t4_30b_mazur_n1():
    m = mazur(1, 3)
    L = m.layers
    L[0].w[:] = np.array([0.15, 0.2, 0.25, 0.3, 0.2, 0.15], np.float32).reshape(3, 2)
    L[0].b[:] = 0.35
    L[2].w[:] = np.array([0.4, 0.45, 0.5, 0.55, 0.5, 0.45], np.float32).reshape(2, 3)
    L[2].b[:] = 0.6
    m.forward(np.array([0.05, 0.1], np.float32))
    close(L[1].data, [0.3775, 0.3925, 0.3750])                  # :29
    close(L[1].ex,   [0.2413, 0.2406, 0.2414])                  # :30  s(1-s)
    close(L[2].data, [0.5933, 0.5969, 0.5927])                  # :31
    close(L[3].data, [1.4022, 1.4914])                          # :32
    close(L[3].ex,   [0.1585, 0.1500])                          # :33
    close(L[4].data, [0.8025, 0.8163])                          # :34
    tgt = np.array([0.01, 0.99], np.float32)
    close(m.loss(orc.LOSS_MSE, tgt), 0.658292, 2e-6)            # :39
    m.backprop(tgt)
    close(L[4].data, [0.7925, -0.1737])                         # :43
    close(L[3].data, [0.7925, -0.1737])                         # :44  sigmoid bwd = identity
    close(L[2].db,   [0.7925, -0.1737])                         # :45
    close(L[2].dw,   [[0.4702, 0.4731, 0.4697], [-0.1031, -0.1037, -0.1029]])   # :46-48
    close(L[2].data, [0.2215, 0.2698, 0.3181])                  # :49
    close(L[1].data, [0.2215, 0.2698, 0.3181])                  # :50
    close(L[0].db,   [0.2215, 0.2698, 0.3181])                  # :51
    close(L[0].dw,   [[0.0111, 0.0221], [0.0135, 0.0270], [0.0159, 0.0318]])    # :52-53
    close(L[0].data, [0.1643, 0.1729])                          # :54-56
    m.sgd(0.5, 0.0)
    close(L[2].w, [[0.1649, 0.2135, 0.2651], [0.6015, 0.5518, 0.5015]])         # :59-61
    close(L[2].b, [0.2037, 0.6869])                             # :63-64
    assert not L[2].dw.any() and not L[2].db.any()              # :62,:65 zero after update
    close(L[0].w, [[0.1445, 0.1889], [0.2433, 0.2865], [0.1920, 0.1341]])       # :66-69
    close(L[0].b, [0.2393, 0.2151, 0.1909])                     # :70-72


# ------------------------------------------------------------------ t4_30c.4th
def test_t4_30c_mazur_n3():
    m = mazur(3, 2)
    L = m.layers
    L[0].w[:] = np.array([0.15, 0.2, 0.25, 0.3], np.float32).reshape(2, 2)
    L[0].b[:] = 0.35
    L[2].w[:] = np.array([0.4, 0.45, 0.5, 0.55], np.float32).reshape(2, 2)
    L[2].b[:] = 0.6
    m.forward(np.array([0.05, 0.1] * 3, np.float32))
    close(L[1].data, [0.3775, 0.3925] * 3)                      # :27
    close(L[1].ex,   [0.2413, 0.2406] * 3)                      # :28
    close(L[2].data, [0.5933, 0.5969] * 3)                      # :29
    close(L[3].data, [1.1059, 1.2249] * 3)                      # :30
    close(L[3].ex,   [0.1868, 0.1755] * 3)                      # :31
    close(L[4].data, [0.7514, 0.7729] * 3)                      # :32
    tgt = np.array([0.01, 0.99] * 3, np.float32)
    close(m.loss(orc.LOSS_MSE, tgt), 0.596742, 2e-6)            # :38
    m.backprop(tgt)
    close(L[4].data, [0.7414, -0.2171] * 3, 1.1e-4)             # :42 (script comment rounds -0.21707 to -0.2172)
    close(L[2].db,   [2.2241, -0.6512], 1.1e-4)                 # :44
    close(L[2].dw,   [[1.3195, 1.3275], [-0.3864, -0.3887]], 1.1e-4)   # :45-47 (comment's -0.3836 is a typo of 3·(-0.2171)·0.5933)
    close(L[1].data, [0.1880, 0.2142] * 3)                      # :48
    close(L[0].db,   [0.5640, 0.6427])                          # :50 verify
    close(L[0].dw,   [[0.0282, 0.0564], [0.0321, 0.0643]])      # :52 verify
    close(L[0].data, [0.0818, 0.1019] * 3)                      # :54 verify
    m.sgd(0.5, 0.0)
    close(L[0].w, [[0.1359, 0.1718], [0.2339, 0.2679]])         # :67 verify
    close(L[0].b, [0.0680, 0.0287])                             # :70 verify


def test_dataset_load_known_values():
    """Dataset::_load (src/mu/dataset.cu:139-143) with the default scale 1/256 (dataset.h:36) and with `128 128 normalize`
    (examples/t4_40b.4th:52: [0,255] -> [-1, 1))"""
    u8 = np.array([0, 1, 127, 128, 255], np.uint8)
    np.testing.assert_array_equal(orc.dataset_load(u8), np.array([0, 1, 127, 128, 255], np.float32) / 256)
    mean, scale = orc.dataset_normalize(128.0, 128.0)
    assert (mean, scale) == (128.0, 1.0 / 128.0)
    np.testing.assert_array_equal(orc.dataset_load(u8, mean, scale), np.array([-1.0, -0.9921875, -0.0078125, 0.0, 0.9921875], np.float32))
    assert orc.dataset_normalize(0.0, 0.0)[1] == 1.0                    # "scale == 0?" -> 1.0
