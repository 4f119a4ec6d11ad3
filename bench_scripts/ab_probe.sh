#!/bin/bash
# development aid: GPU tests + the MNIST step bench (+ optional ncu capture: AB_NCU=<kernel regex> AB_NCU_WHAT=<perf_probe section>)
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu ${AB_K:+-k "$AB_K"} > gpurun_out/ab_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/ab_pytest.log
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
if [ -n "$AB_NCU" ]; then
  timeout 240 ncu --set full --clock-control none --import-source on -k regex:$AB_NCU -s ${AB_NCU_SKIP:-3} -c 1 -o gpurun_out/ab_ncu -f python tests/perf_probe.py ${AB_NCU_WHAT:-cpr} > gpurun_out/ab_ncu.log 2>&1; tail -1 gpurun_out/ab_ncu.log
fi
if [ -n "$AB_ENV2" ]; then run alt "$AB_ENV2"; fi
