"""
refkern.py — talks to oracle/_ref/refkern (the REFERENCE's own CUDA kernels behind a
file protocol, see oracle/ref/refkern.cu).  TEST INFRASTRUCTURE ONLY.

Needs a GPU; used on the B200 box (a) by tests/golden/make_golden.py to generate the
committed fixtures and (b) by the `-m gpu` parity tests when the prebuilt binary travelled.
"""
import os
import struct
import subprocess
import tempfile
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
BIN = os.path.join(_HERE, "_ref", "refkern")
TEN4 = os.path.join(_HERE, "_ref", "ten4")


def available():
    return os.path.exists(BIN) and os.access(BIN, os.X_OK)


def _pack(op, ints=(), flts=(), arrs=()):
    b = op.encode().ljust(16, b"\0")[:16]
    b += struct.pack("<i", len(ints)) + struct.pack("<%di" % len(ints), *[int(x) for x in ints])
    b += struct.pack("<i", len(flts)) + struct.pack("<%df" % len(flts), *[float(x) for x in flts])
    b += struct.pack("<i", len(arrs))
    for a in arrs:
        a = np.ascontiguousarray(a, dtype=np.float32)
        b += struct.pack("<q", a.size) + a.tobytes()
    return b


def run(records, timeout=600):
    """records: list of (op, ints, flts, arrays) → list of lists of flat float32 arrays"""
    with tempfile.TemporaryDirectory() as td:
        req, rsp = os.path.join(td, "req.bin"), os.path.join(td, "rsp.bin")
        with open(req, "wb") as f:
            for r in records:
                f.write(_pack(*r))
        p = subprocess.run([BIN, req, rsp], capture_output=True, text=True, timeout=timeout)
        if p.returncode != 0:
            raise RuntimeError("refkern rc=%d: %s %s" % (p.returncode, p.stdout[-2000:], p.stderr[-2000:]))
        out = []
        with open(rsp, "rb") as f:
            buf = f.read()
        off = 0
        for _ in records:
            (na,) = struct.unpack_from("<i", buf, off); off += 4
            arrs = []
            for _k in range(na):
                (ln,) = struct.unpack_from("<q", buf, off); off += 8
                arrs.append(np.frombuffer(buf, np.float32, ln, off).copy()); off += 4 * ln
            out.append(arrs)
        return out


def one(op, ints=(), flts=(), arrs=()):
    return run([(op, ints, flts, arrs)])[0]
