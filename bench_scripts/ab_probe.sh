#!/bin/bash
# A/B of tuning knobs on the MNIST step (development aid): T4K_CPR2_BWD_THREADS, T4K_SIMT_SPLIT_MULT, T4K_GEMM_MMA
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -x -q -k "cpr or conv_pool or block or fused or train_steps or step_graph or golden or gemm_engines" > gpurun_out/ab_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/ab_pytest.log
for thr in 224 128; do
  echo "== cpr bwd threads=$thr"; T4K_CPR2_BWD_THREADS=$thr timeout 120 python tests/perf_probe.py cpr 2>&1 | tail -1
done
run() { echo "== bench $1"; env $2 timeout 200 python bench.py --no-extras --no-cpu-baseline > gpurun_out/ab_bench_$1.json 2> gpurun_out/ab_bench_$1.err; python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/ab_bench_$1.json").read().strip().splitlines()[-1])
    print("$1", "us/step %.2f" % (d["ms_per_step"] * 1e3), "e2e %.3f M/s" % (d["e2e"]["value"] / 1e6), "loss", d.get("final_loss"))
    for c in d.get("calls", []): print("    %-55s %d %6.2f us" % (c["call"][:55], c["launches"], c["us"]))
except Exception as e:
    print("$1 failed", e)
PY
}
run default "T4K_X=1"
run cpr224 "T4K_CPR2_BWD_THREADS=224"
run mult3 "T4K_SIMT_SPLIT_MULT=3"
run mult4 "T4K_SIMT_SPLIT_MULT=4"
