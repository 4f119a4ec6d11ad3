"""helpers for the -m gpu parity tests: device buffers via torch (plumbing only), calls via the C-ABI"""
import ctypes as C
import numpy as np
import torch

from tensorforth_b200 import lib as t4

L = None


def lib():
    global L
    if L is None:
        L = t4.load()
    return L


_KEEP = []          # keeps device buffers alive while asynchronous kernels use their raw pointers


def dev(a, dtype=np.float32):
    """numpy → CUDA tensor (contiguous); kept alive until release() (the C-ABI sees raw pointers only)"""
    t = torch.from_numpy(np.ascontiguousarray(a, dtype=dtype)).cuda()
    _KEEP.append(t)
    return t


def release():
    torch.cuda.synchronize()
    _KEEP.clear()


def zeros(*shape):
    t = torch.zeros(*shape, dtype=torch.float32, device="cuda")
    _KEEP.append(t)
    return t


def ptr(t, off=0):
    return C.c_void_p(t.data_ptr() + 4 * off)


def host(t):
    torch.cuda.synchronize()
    return t.detach().cpu().numpy()


def ok(rc, what=""):
    t4.check(rc, what)


def assert_close(got, ref, rtol=1e-4, atol=None, what=""):
    """|got-ref| <= rtol*|ref| + atol, atol defaulting to rtol * rms(ref) (FP32 bar of the north star: 1e-4 rel)"""
    got = np.asarray(got, np.float64).ravel()
    ref = np.asarray(ref, np.float64).ravel()
    assert got.shape == ref.shape, (got.shape, ref.shape)
    if atol is None:
        atol = rtol * (np.sqrt(np.mean(ref * ref)) + 1e-30)
    err = np.abs(got - ref) - (rtol * np.abs(ref) + atol)
    bad = int((err > 0).sum())
    assert bad == 0, "%s: %d/%d out of tolerance, max abs err %.3e (ref rms %.3e)" % (
        what, bad, ref.size, np.abs(got - ref).max(), np.sqrt(np.mean(ref * ref)))


def assert_exact(got, ref, what=""):
    got = np.asarray(got); ref = np.asarray(ref)
    assert got.shape == ref.shape or got.size == ref.size
    assert np.array_equal(got.ravel().view(np.uint32), np.asarray(ref, np.float32).ravel().view(np.uint32)), what
