"""integration/arena_shim.cpp (the VM's object store beyond 2 GiB) is host-only code: its allocator — the reference's TLSF
class interface with 64-bit host-side bookkeeping — is stress-tested here on the CPU (tests/arena_driver.cpp): 256-byte alignment, the float of
slack behind every block, no overlap under 20 000 random alloc/free, full coalescing, refused (not wrapped) requests, and offsets beyond 32 bits
in a 24 GiB store.  Needs the reference's header src/mu/tlsf.h (build container only: the driver compiles against it where it lies); the
device-facing half (managed range, preferred location, a 3.3 GB model from Forth) is tests/test_gpu_side_by_side.py on the GPU box."""
import os
import subprocess
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/src"


def test_object_store_allocator_stress_on_cpu():
    if not os.path.exists(os.path.join(REF, "mu", "tlsf.h")):
        pytest.skip("reference sources not present (GPU box): the allocator is compiled against src/mu/tlsf.h")
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    with tempfile.TemporaryDirectory() as d:
        cx = ["g++", "-std=c++17", "-O2", "-I" + REF, "-I" + os.path.join(cuda, "include"), "-w"]
        subprocess.check_call(cx + ["-c", os.path.join(ROOT, "integration", "arena_shim.cpp"), "-o", os.path.join(d, "arena.o")])
        subprocess.check_call(cx + ["-c", os.path.join(ROOT, "tests", "arena_driver.cpp"), "-o", os.path.join(d, "drv.o")])
        exe = os.path.join(d, "arena_test")
        subprocess.check_call(["g++", os.path.join(d, "drv.o"), os.path.join(d, "arena.o"), "-L" + os.path.join(cuda, "lib64"), "-lcudart",
                               "-Wl,-rpath," + os.path.join(cuda, "lib64"), "-o", exe])
        p = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert p.returncode == 0 and p.stdout.strip().endswith("ARENA OK"), p.stdout[-600:] + p.stderr[-300:]
    assert "FAIL" not in p.stdout
