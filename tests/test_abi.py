"""
CPU-side boundary checks (no GPU): libt4k.so builds for sm_100a, loads, and exports every symbol
include/t4k.h declares; the ctypes prototypes cover exactly that set; nothing in the product
package imports the oracle.
"""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "t4k.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(t4k_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_every_declared_symbol():
    import __graft_entry__ as g
    g.build()
    from tensorforth_b200 import lib
    L = lib.load()
    syms = header_symbols()
    assert len(syms) >= 40
    for s in syms:
        assert hasattr(L, s), "libt4k.so does not export %s" % s
    assert sorted(lib.PROTOTYPES) == syms, set(lib.PROTOTYPES) ^ set(syms)
    assert L.t4k_version() == 100
    assert L.t4k_strerror(-1).decode().startswith("t4k")


def test_sass_is_blackwell_native():
    """tcgen05.mma → UTC*MMA, tcgen05.ld → LDTM, cp.async.bulk → UBLKCP (B200_PROFILING.md)"""
    so = os.path.join(ROOT, "tensorforth_b200", "libt4k.so")
    sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
    assert "UTCHMMA" in sass or "UTCMMA" in sass
    assert "LDTM" in sass and "UBLKCP" in sass
    assert "sm_100a" in subprocess.run(["cuobjdump", "-lelf", so], capture_output=True, text=True).stdout


def test_product_never_imports_the_oracle():
    bad = []
    for d, _, files in os.walk(os.path.join(ROOT, "tensorforth_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".cc")):
                txt = open(os.path.join(d, f), errors="ignore").read()
                if re.search(r"(^|\s)(from|import)\s+oracle|t4_oracle\.h|libt4oracle", txt):
                    bad.append(f)
    assert not bad, bad


def test_no_gpu_means_error_not_fallback():
    import torch
    if torch.cuda.is_available():
        return
    from tensorforth_b200 import lib
    L = lib.load()
    assert L.t4k_device_count() == 0
    import ctypes as C
    buf = (C.c_float * 16)()
    rc = L.t4k_map(lib.FILL, C.cast(buf, C.c_void_p), 1.0, 16, None)
    assert rc != 0 and buf[0] == 0.0           # launch fails loudly; host memory is never touched by a CPU path


def test_host_library_exports_every_symbol_of_t4host_h():
    """libt4host.so (the Tensor/Model/Dataset class mirror) must export what include/t4host.h declares, and the ctypes face
    must bind exactly those it uses"""
    import ctypes as C
    hdr = open(os.path.join(ROOT, "include", "t4host.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    syms = sorted(set(re.findall(r"\b(t4h_[a-z0-9_]+)\s*\(", hdr)))
    assert len(syms) >= 50, len(syms)
    L = C.CDLL(os.path.join(ROOT, "tensorforth_b200", "libt4host.so"))
    missing = [s for s in syms if not hasattr(L, s)]
    assert not missing, missing
    from tensorforth_b200 import host
    unknown = [s for s in host.PROTOTYPES if s not in syms]
    assert not unknown, unknown


def test_integration_index_matches_header():
    """INTEGRATION.md §8 lists every function include/t4k.h declares, with the reference lines the header cites for it
    (regenerate with `python bench_scripts/abi_index.py --write`); only library plumbing may be without a reference counterpart"""
    import importlib.util
    spec = importlib.util.spec_from_file_location("abi_index", os.path.join(ROOT, "bench_scripts", "abi_index.py"))
    ai = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ai)
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    block = doc[doc.index(ai.BEGIN) + len(ai.BEGIN):doc.index(ai.END)].strip()
    assert block == ai.table().strip(), "INTEGRATION.md §8 is stale: python bench_scripts/abi_index.py --write"
    ents = ai.entries()
    funcs = [s for s in header_symbols() if s not in ("t4k_comm",)]
    listed = {e[0] for e in ents}
    missing = [s for s in funcs if s not in listed and not s.endswith("_t")]
    assert not missing, missing
    uncited = sorted(e[0] for e in ents if not e[2])
    plumbing = {"t4k_version", "t4k_strerror", "t4k_device_count", "t4k_sm_count", "t4k_sync", "t4k_launch_count", "t4k_set_workspace_bank",
                "t4k_set_carveout", "t4k_set_pdl", "t4k_set_conv_engine"}
    assert set(uncited) <= plumbing, set(uncited) - plumbing


def test_host_mirror_fails_loudly_without_a_gpu():
    """the Tensor / Model mirror (libt4host.so) has no CPU path either: init raises, nothing is computed on the host"""
    import torch
    if torch.cuda.is_available():
        return
    from tensorforth_b200 import host as th
    from tensorforth_b200 import lib
    import pytest
    with pytest.raises(lib.T4KError) as e:
        th.init(0)
    assert "no CUDA device" in str(e.value) and "no CPU fallback" in str(e.value)
