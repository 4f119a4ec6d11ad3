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
