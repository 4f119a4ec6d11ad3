"""The reference's own Forth VM, MMU, model.cpp, loss.cpp, gradient.cu and printer — unmodified, compiled where they
lie — linked on libt4k.so (integration/_build/ten4_b200) must print what the reference build (oracle/_ref/ten4) prints
for the same Forth text: the reference's example scripts (run unchanged) and the deterministic training scripts under
integration/scripts/.  Both binaries are built in the build container (they need the reference sources) and travel to
the GPU box as git-ignored artefacts; without them the test is skipped."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "ten4")
NEW = os.path.join(ROOT, "integration", "_build", "ten4_b200")


@pytest.mark.gpu
def test_reference_vm_on_libt4k_prints_what_the_reference_prints(tmp_path):
    if not (os.path.exists(REF) and os.path.exists(NEW)):
        pytest.skip("oracle/_ref/ten4 or integration/_build/ten4_b200 not built")
    out = os.path.join(ROOT, "gpurun_out", "side_by_side") if os.path.isdir(os.path.join(ROOT, "gpurun_out")) else str(tmp_path / "sbs")
    p = subprocess.run(["bash", os.path.join(ROOT, "integration", "run_side_by_side.sh"), out], capture_output=True, text=True, timeout=900)
    summary = open(os.path.join(out, "summary.txt")).read() if os.path.exists(os.path.join(out, "summary.txt")) else p.stdout
    sys.stdout.write(summary)
    assert " OK " in summary, p.stdout[-2000:] + p.stderr[-2000:]
    assert "DIFF" not in summary and "MISSING" not in summary, summary


@pytest.mark.gpu
def test_object_store_beyond_2gib_from_forth():
    """`ten4_b200` carries integration/arena_shim (object store sized for the device, 64-bit host-side allocator,
    device-preferred managed memory): a 4096^2 `@` and an N=1024 64-channel 56x56 conv2d layer (3.3 GB of tensors) run from Forth text on the
    new kernels.  The reference's own 2 GiB TLSF store (src/ten4_config.h:67, src/mu/tlsf.h:19-31) cannot hold the model."""
    if not os.path.exists(NEW):
        pytest.skip("integration/_build/ten4_b200 not built")
    src = open(os.path.join(ROOT, "integration", "scripts_b200", "big_arena.4th")).read()
    p = subprocess.run([NEW], input=src, capture_output=True, text=True, timeout=600, cwd=ROOT)
    out = p.stdout
    sys.stdout.write(out[-1500:])

    def val(tag):
        import re
        m = re.search(re.escape(tag) + r"\s*(-?[0-9.]+(?:e[-+]?\d+)?)", out)
        assert m, (tag, out[-1500:], p.stderr[-500:])
        return float(m.group(1))
    assert abs(val("gemm sum/4096^3=") - 1.0) < 1e-3
    assert abs(val("conv out max=") - 9 * 64 * 0.5 * 0.001) < 1e-4          # interior pixel: all nine taps in the image
    assert abs(val("conv out min=") - 4 * 64 * 0.5 * 0.001) < 1e-4          # corner pixel: four taps
    assert val("dw sum/1e6=") > 0


@pytest.mark.gpu
def test_dconv2d_word_through_the_reference_vm():
    """the conv-transpose layer driven from Forth text: the reference's VM and Model::add (unmodified) on integration/model_shim.cu's L_DCONV cases
    (t4k_dconv2d_fwd / _bwd); closed-form values for constant tensors (see the script's header)"""
    if not os.path.exists(NEW):
        pytest.skip("integration/_build/ten4_b200 not built")
    src = open(os.path.join(ROOT, "integration", "scripts_b200", "dconv2d.4th")).read()
    p = subprocess.run([NEW], input=src, capture_output=True, text=True, timeout=300, cwd=ROOT)
    out = p.stdout
    sys.stdout.write(out[-800:])
    import re

    def val(tag):
        m = re.search(re.escape(tag) + r"\s*(-?[0-9.]+(?:e[-+]?\d+)?)", out)
        assert m, (tag, out[-800:], p.stderr[-300:])
        return float(m.group(1))
    assert abs(val("out max=") - 0.16) < 1e-4 and abs(val("out min=") - 0.04) < 1e-4
    # every input element reaches 16 taps (4 x 4), minus those that fall outside the 8 x 8 output: sum over the output = 0.005 * 6 * sum over inputs of its taps inside
    assert abs(val("out sum=") - 2 * 6 * 8 * 0.005 * (4 * 9 + 8 * 12 + 4 * 16)) < 1e-2
    assert abs(val("dx max=") - 16 * 6 * 0.01) < 1e-4
    assert abs(val("db sum=") - 2 * 8 * 8 * 6) < 1e-2
