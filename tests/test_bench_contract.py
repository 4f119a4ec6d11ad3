"""bench.py contract checks that need no GPU: the reference arm's JSON line (here the reference binary cannot start — no CUDA driver — so the
arm falls back to the CPU oracle port, the same code path a box without oracle/_ref/ten4 takes), rank handling under torchrun, and that both
arms name the same workload / metric (`same_config` was false because of a string mismatch)."""
import inspect
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REQUIRED = ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
            "config", "cpu_baseline", "e2e")


def run_bench(args, env_extra=None):
    env = dict(os.environ)
    env.update(env_extra or {})
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)
    return p


def test_reference_arm_prints_one_contract_line_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        return                                              # on a GPU box the arm runs the reference binary: covered by the driver
    p = run_bench(["--impl", "reference", "--steps", "2", "--warmup", "3"])
    assert p.returncode == 0, p.stderr[-500:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines                           # ONE JSON line on stdout, everything else on stderr
    d = json.loads(lines[0])
    for k in REQUIRED:
        assert k in d, k
    assert d["impl"] == "reference" and d["metric"] == "mnist_cnn_train_samples_per_sec" and d["unit"] == "samples/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["dtype"] == "f32"
    assert d["value"] > 0 and abs(d["ms_per_step"] - 512 / d["value"] * 1e3) < 1e-6 * d["ms_per_step"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_other_ranks_exit_silently():
    p = run_bench(["--impl", "reference", "--gpus", "2", "--steps", "2"], {"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"})
    assert p.returncode == 0 and p.stdout.strip() == "", (p.returncode, p.stdout[-200:])


def test_both_arms_name_the_same_workload_and_metric():
    sys.path.insert(0, ROOT)
    import bench
    src = inspect.getsource(bench)
    assert src.count('"workload": WORKLOAD') == 2                                   # reference_arm and main
    assert src.count('"metric": "mnist_cnn_train_samples_per_sec"') == 2
    assert "N=512 per GPU" in bench.WORKLOAD and "nn.adam" in bench.WORKLOAD
    # the Forth text of the reference arm is the same step: forward + loss.ce + backprop + nn.adam on the t4_40a CNN at the same batch
    txt = bench.ref_script("mnist", 3, 20, bench.BATCH)
    for w in ("0 trace", "512 constant N", "0.5 10 conv2d 2 maxpool relu flatten 100 linear relu 10 linear softmax", "forward", "loss.ce", "backprop", "0.001 nn.adam"):
        assert w in txt, w


def test_our_arm_refuses_to_run_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        return
    p = run_bench(["--steps", "2", "--no-extras", "--no-cpu-baseline"])
    assert p.returncode != 0 and p.stdout.strip() == ""                             # no CPU path, no fabricated line
    assert "needs a CUDA device" in p.stderr
