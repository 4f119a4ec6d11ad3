"""
host.py — Python face of libt4host.so (include/t4host.h): `Tensor` and `Model` with the reference's
method names (src/mu/tensor.h, src/nn/model.h) and the Forth words' semantics (src/vm/tenvm.cpp,
netvm.cpp).  Pure plumbing: every method is one C call; the math runs in libt4k.so on the GPU.
Raises T4KError when the CUDA libraries are missing — there is no CPU path.
"""
import ctypes as C
import os
import numpy as np

from . import lib as _k
from .lib import T4KError

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "libt4host.so")
_p, _i, _f, _l, _u = C.c_void_p, C.c_int, C.c_float, C.c_int64, C.c_uint32
_fp = C.POINTER(C.c_float)

PROTOTYPES = {
    "t4h_last_error": (C.c_char_p, []), "t4h_init": (_i, [_i]), "t4h_stream": (_p, []), "t4h_sync": (_i, []),
    "t4h_launch_count": (C.c_long, []),
    "t4h_tensor_new": (_p, [_i, _u, _u, _u, _u]), "t4h_tensor_free": (None, [_p]), "t4h_tensor_data": (_p, [_p]),
    "t4h_tensor_shape": (_i, [_p, C.POINTER(_u)]), "t4h_tensor_numel": (_l, [_p]), "t4h_tensor_rank": (_i, [_p]),
    "t4h_tensor_h2d": (_i, [_p, _p, _l]), "t4h_tensor_d2h": (_i, [_p, _p, _l]),
    "t4h_tensor_reshape": (_i, [_p, _i, _u, _u, _u, _u]), "t4h_tensor_copy": (_p, [_p]),
    "t4h_tensor_map": (_i, [_p, _i, _f]), "t4h_tensor_identity": (_i, [_p]), "t4h_tensor_rand": (_i, [_p, _i]),
    "t4h_ten_op_s": (_i, [_i, _p, _f, _p]), "t4h_ten_op_t": (_i, [_i, _p, _p, _p]),
    "t4h_mm": (_i, [_p, _p, _p, _i, _i, _i]), "t4h_gemm": (_i, [_i, _p, _p, _p, _f, _f, _i, _i]),
    "t4h_matmul": (_p, [_p, _p]), "t4h_transpose": (_p, [_p]),
    "t4h_tensor_sum": (_f, [_p]), "t4h_tensor_avg": (_f, [_p]), "t4h_tensor_std": (_f, [_p]), "t4h_tensor_norm": (_f, [_p]),
    "t4h_tensor_max": (_f, [_p]), "t4h_tensor_min": (_f, [_p]), "t4h_tensor_dot": (_f, [_p, _p]),
    "t4h_tensor_loss": (_f, [_p, _i, _p]),
    "t4h_model_new": (_p, [_u, _u, _u, _u]), "t4h_model_free": (None, [_p]),
    "t4h_model_add": (_i, [_p, _i, _u, _f, C.POINTER(C.c_uint16)]), "t4h_model_numel": (_i, [_p]),
    "t4h_model_layer": (_p, [_p, _i]), "t4h_model_param": (_p, [_p, _i, _i]), "t4h_model_set_param": (_i, [_p, _i, _i, _p]),
    "t4h_model_train": (_i, [_p, _i]), "t4h_model_fuse": (_i, [_p, _i]), "t4h_model_forward": (_i, [_p, _p]), "t4h_model_backprop": (_i, [_p, _p]),
    "t4h_model_loss": (_f, [_p, _i, _p]), "t4h_model_loss_async": (_i, [_p, _i, _p, _p]),
    "t4h_model_onehot_labels": (_i, [_p, _p]), "t4h_model_onehot_set": (_i, [_p, _p]), "t4h_model_hit": (_i, [_p, _i]),
    "t4h_model_sgd": (_i, [_p, _f, _f]), "t4h_model_adam": (_i, [_p, _f, _f, _f]), "t4h_model_adamw": (_i, [_p, _f, _f, _f, _f]),
    "t4h_model_arena": (_i, [_p, C.POINTER(_p), C.POINTER(_p), C.POINTER(_l)]),
    "t4h_model_step_graph": (_i, [_p, _p, _p, _i, _p, _i, _f, _f, _f, _f]),
    "t4h_model_dp_attach": (_i, [_p, _p, _p, _i]),
    "t4h_model_dp_shard": (_i, [_p, _i, _i, _p]),
    "t4h_model_bn_channels": (_i, [_p]),
    "t4h_tensor_rand_sharded": (_i, [_p, _i, _i, _i]),
    "t4h_use_lane": (_i, [_i]),
    "t4h_set_dp_early": (_i, [_i]),
    "t4h_set_dp_rest": (_i, [_i]),
    "t4h_side_stream": (_p, []),
    "t4h_capture_begin": (_i, []), "t4h_capture_end": (_p, []), "t4h_graph_launch": (_i, [_p]), "t4h_graph_free": (None, [_p]),
    "t4h_model_save_state": (_i, [_p, C.c_char_p]),
    "t4h_model_save": (_i, [_p, C.c_char_p]), "t4h_model_load": (_i, [_p, C.c_char_p]),
    "t4h_dataset_create": (_p, [_i, _i, _i, _i]), "t4h_dataset_destroy": (None, [_p]), "t4h_dataset_normalize": (None, [_p, _f, _f]),
    "t4h_dataset_stage": (_i, [_p, _p, _p, _i]), "t4h_dataset_commit": (_i, [_p]), "t4h_dataset_tensor": (_p, [_p]),
    "t4h_dataset_labels": (_p, [_p]), "t4h_model_forward_ds": (_i, [_p, _p]),
    "t4h_model_step_graph_ds": (_i, [_p, _p, _i, _p, _i, _f, _f, _f, _f]),
    "t4h_model_train_step_ds": (_i, [_p, _p, _i, _p, _i, _f, _f, _f, _f, C.POINTER(_f)]), "t4h_model_train_flush": (_i, [_p, C.POINTER(_f)]),
}
_lib = None


def load():
    global _lib
    if _lib is None:
        _k.load()                       # libt4k.so first (dependency) — raises if the CUDA extension is missing
        if not os.path.exists(SO_PATH):
            raise T4KError("libt4host.so not built (%s)" % SO_PATH)
        L = C.CDLL(SO_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def _err():
    return load().t4h_last_error().decode(errors="replace")


def init(device=0):
    rc = load().t4h_init(device)
    if rc:
        raise T4KError("t4h_init(%d): %s" % (device, _err()))


def use_lane(k):
    """tests: switch the process to stream set k (several data-parallel ranks of one process, each on its own lane)"""
    _k.check(load().t4h_use_lane(int(k)), "use_lane")


def sync():
    load().t4h_sync()


def stream():
    return load().t4h_stream()


class Tensor:
    """src/mu/tensor.h:51-190 — shape is (N,H,W,C); rank 1 vectors are H=numel (tensor.cu:461-484)"""

    def __init__(self, handle, owned=True):
        if not handle:
            raise T4KError("tensor: " + _err())
        self.h, self.owned = handle, owned

    # ---- words `vector matrix tensor vector{ matrix{` (tenvm.cpp:458-474)
    @staticmethod
    def vector(n, values=None):
        t = Tensor(load().t4h_tensor_new(1, 1, n, 1, 1)); return t.set(values) if values is not None else t

    @staticmethod
    def matrix(h, w, values=None):
        t = Tensor(load().t4h_tensor_new(2, 1, h, w, 1)); return t.set(values) if values is not None else t

    @staticmethod
    def tensor(n, h, w, c, values=None):
        t = Tensor(load().t4h_tensor_new(4, n, h, w, c)); return t.set(values) if values is not None else t

    @staticmethod
    def from_numpy(a):
        a = np.ascontiguousarray(a, np.float32)
        if a.ndim == 1: return Tensor.vector(a.size, a)
        if a.ndim == 2: return Tensor.matrix(a.shape[0], a.shape[1], a)
        return Tensor.tensor(*a.shape, values=a)

    def __del__(self):
        try:
            if self.owned and self.h and _lib is not None:
                _lib.t4h_tensor_free(self.h)
        except Exception:
            pass

    def set(self, values):
        a = np.ascontiguousarray(values, np.float32)
        assert a.size == self.numel, (a.size, self.numel)
        load().t4h_tensor_h2d(self.h, a.ctypes.data_as(_p), a.size); sync()        # `a` must outlive the async copy
        return self

    @property
    def numel(self): return int(load().t4h_tensor_numel(self.h))
    @property
    def rank(self): return int(load().t4h_tensor_rank(self.h))
    @property
    def data_ptr(self): return load().t4h_tensor_data(self.h)

    @property
    def shape(self):
        s = (_u * 4)(); load().t4h_tensor_shape(self.h, s); return tuple(int(x) for x in s)     # (N,H,W,C)

    def numpy(self):
        s = self.shape
        out = np.empty(self.numel, np.float32)
        load().t4h_tensor_d2h(self.h, out.ctypes.data_as(_p), out.size)
        return out.reshape(s if self.rank == 4 else (s[1], s[2]) if self.rank == 2 else (self.numel,))

    def reshape(self, *dims):           # reshape2 / reshape4 / flatten words
        r = {1: (1, 1, dims[0], 1, 1), 2: (2, 1) + tuple(dims) + (1,), 4: (4,) + tuple(dims)}[len(dims)]
        if load().t4h_tensor_reshape(self.h, *r): raise T4KError(_err())
        return self

    def copy(self): return Tensor(load().t4h_tensor_copy(self.h))
    def map(self, op, v=0.0): _k.check(load().t4h_tensor_map(self.h, op, v), "map"); return self
    def fill(self, v): return self.map(_k.FILL, v)
    def zeros(self): return self.fill(0.0)
    def ones(self): return self.fill(1.0)
    def eye(self): load().t4h_tensor_identity(self.h); return self
    def rand(self): _k.check(load().t4h_tensor_rand(self.h, _k.UNIFORM)); return self
    def randn(self): _k.check(load().t4h_tensor_rand(self.h, _k.NORMAL)); return self

    def rand_sharded(self, rank, world, normal=False):
        """this tensor holds shard `rank` of `world` equal shards of a batch-major tensor: draw it as one device holding the whole would"""
        _k.check(load().t4h_tensor_rand_sharded(self.h, _k.NORMAL if normal else _k.UNIFORM, rank, world)); return self

    def _bin(self, op, other, out=None):
        out = out or self
        rc = load().t4h_ten_op_t(op, self.h, other.h, out.h) if isinstance(other, Tensor) else load().t4h_ten_op_s(op, self.h, float(other), out.h)
        if rc: raise T4KError(_err())
        return out

    def __iadd__(self, o): return self._bin(_k.ADD, o)        # += -= *= /= (destructive, as in Forth)
    def __isub__(self, o): return self._bin(_k.SUB, o)
    def __imul__(self, o): return self._bin(_k.MUL, o)
    def __itruediv__(self, o): return self._bin(_k.DIV, o)

    def __matmul__(self, o):            # `@` / `matmul` (TensorVM::_tdot rank rules)
        h = load().t4h_matmul(self.h, o.h)
        if not h: raise T4KError(_err())
        return Tensor(h)

    def transpose(self): return Tensor(load().t4h_transpose(self.h))
    def sum(self): return float(load().t4h_tensor_sum(self.h))
    def avg(self): return float(load().t4h_tensor_avg(self.h))
    def std(self): return float(load().t4h_tensor_std(self.h))
    def norm(self): return float(load().t4h_tensor_norm(self.h))
    def max(self): return float(load().t4h_tensor_max(self.h))
    def min(self): return float(load().t4h_tensor_min(self.h))
    def dot(self, o): return float(load().t4h_tensor_dot(self.h, o.h))
    def loss(self, op, tgt): return float(load().t4h_tensor_loss(self.h, op, tgt.h))


def gemm(variant, A, B, Cm, alpha, beta, tA=False, tB=False):
    """`gemm1..gemm4` words: O = alpha*A@B + beta*C on a hard copy of C (TensorVM::gemm, tenvm.cpp:210-237)"""
    O = Cm.copy()
    load().t4h_gemm(variant, A.h, B.h, O.h, alpha, beta, int(tA), int(tB))
    return O


class Model:
    """src/nn/model.h:36-164 + the NN vocabulary of src/vm/netvm.cpp:291-485"""

    def __init__(self, n, h, w, c):     # `n h w c nn.model`
        init()
        self.h = load().t4h_model_new(n, h, w, c)
        if not self.h: raise T4KError(_err())
        self._keep = []

    def __del__(self):
        try:
            if self.h and _lib is not None: _lib.t4h_model_free(self.h)
        except Exception:
            pass

    def add(self, layer, n=0, bias=0.0, opt=None):
        o = (C.c_uint16 * 4)(*opt) if opt is not None else None
        if load().t4h_model_add(self.h, layer, n, bias, o): raise T4KError("Model#add: " + _err())
        return self

    # word-level sugar with the VM's default parameters (netvm.cpp:20-133, 211-226)
    def conv2d(self, bias, c, k=3, s=1, p=1, d=1): return self.add(_k.L_CONV, c, bias, [k, s, p, d])   # NetVM::_conv defaults (netvm.h:51)
    def conv1x1(self, bias, c): return self.add(_k.L_CONV, c, bias, [1, 1, 0, 1])
    def linear(self, n, bias=1.0): return self.add(_k.L_LINEAR, n, bias)
    def flatten(self): return self.add(_k.L_FLATTEN)
    def relu(self): return self.add(_k.L_RELU)
    def tanh(self): return self.add(_k.L_TANH)
    def sigmoid(self): return self.add(_k.L_SIGMOID)
    def selu(self): return self.add(_k.L_SELU)
    def leakyrelu(self, a=0.01): return self.add(_k.L_LEAKYRL, 0, a)
    def elu(self, a=1.0): return self.add(_k.L_ELU, 0, a)
    def dropout(self, p): return self.add(_k.L_DROPOUT, 0, p)
    def softmax(self): return self.add(_k.L_SOFTMAX)
    def logsoftmax(self): return self.add(_k.L_LOGSMAX)
    def maxpool(self, k): return self.add(_k.L_MAXPOOL, k)
    def avgpool(self, k): return self.add(_k.L_AVGPOOL, k)
    def minpool(self, k): return self.add(_k.L_MINPOOL, k)
    def dconv2d(self, bias, c, k=4, s=2, p=0, d=1): return self.add(_k.L_DCONV, c, bias, [k, s, p, d])   # word `dconv2d` = _conv(4, true, 2): 4x4, stride 2 (netvm.cpp:315)
    def batchnorm(self, m=0.1): return self.add(_k.L_BATCHNM, 0, m)
    def upsample(self, k, m=0.0): return self.add(_k.L_USAMPLE, k, m)

    def __len__(self): return load().t4h_model_numel(self.h)

    def layer(self, i):                 # `n@`
        h = load().t4h_model_layer(self.h, i)
        if not h: raise IndexError(i)
        return Tensor(h, owned=False)

    def _param(self, i, which):
        h = load().t4h_model_param(self.h, i, which)
        return Tensor(h, owned=False) if h else None

    def w(self, i): return self._param(i, 0)     # nn.w
    def b(self, i): return self._param(i, 1)     # nn.b
    def dw(self, i): return self._param(i, 2)    # nn.dw
    def db(self, i): return self._param(i, 3)    # nn.db
    def ex(self, i): return self._param(i, 4)    # nn.ex

    def set_w(self, i, t):              # `nn.w=`
        t = t if isinstance(t, Tensor) else Tensor.from_numpy(np.asarray(t, np.float32).ravel())
        if load().t4h_model_set_param(self.h, i, 0, t.h): raise T4KError(_err())
        return self

    def set_b(self, i, t):              # `nn.b=`
        t = t if isinstance(t, Tensor) else Tensor.from_numpy(np.asarray(t, np.float32).ravel())
        if load().t4h_model_set_param(self.h, i, 1, t.h): raise T4KError(_err())
        return self

    def trainable(self, on): load().t4h_model_train(self.h, int(on)); return self
    def fuse(self, on): load().t4h_model_fuse(self.h, int(on)); return self

    def forward(self, x):
        if load().t4h_model_forward(self.h, x.h): raise T4KError(_err())
        return self

    def backprop(self, tgt=None):
        if load().t4h_model_backprop(self.h, tgt.h if tgt is not None else None): raise T4KError(_err())
        return self

    def loss(self, op, tgt=None): return float(load().t4h_model_loss(self.h, op, tgt.h if tgt is not None else None))
    def loss_async(self, op, tgt, loss_dev_ptr): return load().t4h_model_loss_async(self.h, op, tgt.h, loss_dev_ptr)
    def onehot_labels(self, labels_dev_ptr): load().t4h_model_onehot_labels(self.h, labels_dev_ptr); return self
    def set_onehot(self, hot): self._keep.append(hot); load().t4h_model_onehot_set(self.h, hot.h); return self
    def hit(self, recalc=True): return int(load().t4h_model_hit(self.h, int(recalc)))
    def sgd(self, lr, b=0.9): load().t4h_model_sgd(self.h, lr, b); return self
    def adam(self, lr, b1=0.9, b2=0.999): load().t4h_model_adam(self.h, lr, b1, b2); return self
    def adamw(self, lr, wd=0.001, b1=0.9, b2=0.999): load().t4h_model_adamw(self.h, lr, wd, b1, b2); return self

    def arena(self):
        """(G_ptr, DG_ptr, total floats) of the flat parameter / gradient arenas"""
        g, dg, n = _p(), _p(), _l()
        _k.check(load().t4h_model_arena(self.h, C.byref(g), C.byref(dg), C.byref(n)), "arena")
        return g.value, dg.value, n.value

    def save(self, fname, opt_state=False):              # word `save` ( N adr len -- N ): the reference's model file (src/io/aio_model.cpp)
        if (load().t4h_model_save_state if opt_state else load().t4h_model_save)(self.h, str(fname).encode()): raise T4KError(_err())
        return self

    def load(self, fname):              # word `load`: parameters into an already built model
        if load().t4h_model_load(self.h, str(fname).encode()): raise T4KError(_err())
        return self

    def dp_shard(self, rank, world, comm_stat=None):
        """this model holds shard `rank` of `world` equal shards of the global batch (dropout masks at the shard's global offsets; batch-norm
        statistics summed over the ranks on `comm_stat`, a t4k_comm_t of its own — required before dp_attach for a model with batchnorm)"""
        _k.check(load().t4h_model_dp_shard(self.h, rank, world, comm_stat), "dp_shard")
        return self

    def bn_channels(self):
        return int(load().t4h_model_bn_channels(self.h))

    def dp_attach(self, comm, scal_dev_ptr=None, nscal=0):
        """data parallel: from now on the optimizer calls sum the gradient arena over the ranks of `comm` (a connected
        t4k_comm_t handle, see dp.PeerComm) inside the optimizer kernel; `nscal` device floats ride along (summed)"""
        _k.check(load().t4h_model_dp_attach(self.h, comm, scal_dev_ptr, nscal), "dp_attach")
        return self

    def forward_ds(self, ds):
        """Model::forward(Dataset&): forward + one-hot of the batch labels + hit count, all on the device"""
        if load().t4h_model_forward_ds(self.h, ds.h): raise T4KError(_err())
        return self

    def step_graph_ds(self, ds, loss_op, loss_dev_ptr, optimizer=2, lr=1e-3, b1=0.9, b2=0.999, wd=0.0):
        """one iteration of `ds for forward loss backprop nn.adam next`: commit the staged batch + one-hot + captured train step"""
        return load().t4h_model_step_graph_ds(self.h, ds.h, loss_op, loss_dev_ptr, optimizer, lr, b1, b2, wd)

    def train_step_ds(self, ds, loss_op, loss_dev_ptr, optimizer=2, lr=1e-3, b1=0.9, b2=0.999, wd=0.0):
        """step_graph_ds + pipelined loss read-back: returns the loss of the PREVIOUS call (nan on the first)"""
        prev = _f()
        _k.check(load().t4h_model_train_step_ds(self.h, ds.h, loss_op, loss_dev_ptr, optimizer, lr, b1, b2, wd, C.byref(prev)), "train_step_ds")
        return prev.value

    def train_flush(self):
        last = _f()
        _k.check(load().t4h_model_train_flush(self.h, C.byref(last)), "train_flush")
        return last.value

    def step_graph(self, x, tgt, loss_op, loss_dev_ptr, optimizer=2, lr=1e-3, b1=0.9, b2=0.999, wd=0.0):
        """forward + loss + backprop + optimizer as one replayed CUDA graph (optimizer: 0 sgd, 2 adam, 3 adamw)"""
        return load().t4h_model_step_graph(self.h, x.h, tgt.h, loss_op, loss_dev_ptr, optimizer, lr, b1, b2, wd)


class Dataset:
    """Mini-batch feeder mirroring the reference's Dataset (src/mu/dataset.h): `N dataset <name>` + `normalize`.  The loader
    (file parsing) is the caller's; stage() takes the raw uint8 image block [n,H,W,C] and uint8 labels [n] of a mini-batch
    (numpy arrays or pinned torch uint8 tensors) and copies the BYTES asynchronously; commit() normalises on the device."""

    def __init__(self, N, H, W, C_):
        self.h = load().t4h_dataset_create(N, H, W, C_)
        self.N, self.shape = N, (N, H, W, C_)
        self._keep = []

    def normalize(self, mean, scale):
        load().t4h_dataset_normalize(self.h, mean, scale); return self

    @staticmethod
    def _ptr(a):
        if hasattr(a, "data_ptr"):
            return C.c_void_p(a.data_ptr()), a.numel()
        a = np.ascontiguousarray(a, dtype=np.uint8)
        return C.c_void_p(a.ctypes.data), a.size, a

    def stage(self, images_u8, labels_u8):
        pi, pl = self._ptr(images_u8), self._ptr(labels_u8)
        self._keep = (self._keep + [pi, pl])[-8:]                  # host blocks stay alive while the async copies run
        if load().t4h_dataset_stage(self.h, pi[0], pl[0], pl[1]): raise T4KError(_err())
        return self

    def commit(self):
        if load().t4h_dataset_commit(self.h): raise T4KError(_err())
        return self

    @property
    def tensor(self):
        return Tensor(load().t4h_dataset_tensor(self.h), owned=False)

    @property
    def labels_ptr(self): return C.c_void_p(load().t4h_dataset_labels(self.h))

    def __del__(self):
        try:
            if self.h: load().t4h_dataset_destroy(self.h); self.h = None
        except Exception:
            pass


def mnist_cnn(N):
    """the MNIST CNN of examples/t4_40a.4th:10-13 / t4_30e.4th:49-55 (BASELINE config 3)"""
    return (Model(N, 28, 28, 1).conv2d(0.5, 10).maxpool(2).relu().flatten().linear(100).relu().linear(10).softmax())


def gan_discriminator(N):
    """examples/t4_40b.4th:37-41"""
    return (Model(N, 28, 28, 1).linear(512).leakyrelu(0.2).dropout(0.3).linear(256).leakyrelu(0.2).dropout(0.3).linear(1).sigmoid())


def gan_generator(N):
    """examples/t4_40b.4th:44-48"""
    return (Model(N, 128, 1, 1).linear(256).leakyrelu(0.2).linear(512).leakyrelu(0.2).linear(784).tanh())


class Graph:
    """A sequence of library calls captured once into a CUDA graph and replayed: `g = Graph(lambda: gan_iteration(..., losses=False))`,
    then `g()` per iteration.  The callable must not read anything back on the host; run it eagerly once before capturing."""

    def __init__(self, fn):
        _k.check(load().t4h_capture_begin(), "capture_begin")
        try:
            fn()
        finally:
            self.h = load().t4h_capture_end()
        if not self.h:
            raise T4KError("graph capture: " + _err())

    def __call__(self):
        _k.check(load().t4h_graph_launch(self.h), "graph_launch")

    def __del__(self):
        try:
            if getattr(self, "h", None): load().t4h_graph_free(self.h); self.h = None
        except Exception:
            pass


def gan_discriminator(N, p=0.3):
    """examples/t4_40b.4th:37-41"""
    return (Model(N, 28, 28, 1).linear(512).leakyrelu(0.2).dropout(p).linear(256).leakyrelu(0.2).dropout(p).linear(1).sigmoid())


def gan_iteration(D, G, real, z_d, z_g, REAL, FAKE, d_lr=1e-4, g_lr=4e-4, b1=0.5, losses=True):
    """one `train_d train_g` of examples/t4_40b.4th:60-67.  real: [N,28,28,1]; z_d / z_g: the latent batches the two `F` calls
    draw ([N,128,1,1]); REAL / FAKE: the [N,1,1,1] ones / zeros targets.  Returns (loss_dr, loss_df, loss_gr) like the script's
    `_dr _df _gr` (host reads: each synchronises) or None with losses=False."""
    BCE = _k.LOSS_BCE
    out = []
    D.trainable(1)                                                      # train_d
    D.forward(real)
    if losses: out.append(D.loss(BCE, REAL))
    D.backprop(REAL)
    G.forward(z_d); D.forward(G.layer(-1))                              # F: G's output (same numel as [N,28,28,1]) feeds D
    if losses: out.append(D.loss(BCE, FAKE))
    D.backprop(FAKE)
    D.adam(d_lr, b1)
    D.trainable(0)                                                      # train_g: D passes the gradient through, no dW/dB
    G.forward(z_g); D.forward(G.layer(-1))
    if losses: out.append(D.loss(BCE, REAL))
    D.backprop(REAL)
    G.backprop(D.layer(0))                                              # `0 n@ G swap backprop`: dX of D's input is G's output gradient
    G.adam(g_lr, b1)
    return tuple(out) if losses else None
