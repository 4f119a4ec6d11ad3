#!/bin/bash
# round-end evidence run (1 GPU): full GPU test suite, bench lines (ours + reference arm), launch list and ncu --set full of one eager step
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/final_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/final_pytest.log
timeout 300 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference > gpurun_out/final_bench_ref.json 2> gpurun_out/final_bench_ref.err; echo "ref rc=$?"
T4K_PDL=1 timeout 200 python bench.py --no-extras --no-cpu-baseline > gpurun_out/final_bench_pdl.json 2> gpurun_out/final_bench_pdl.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 2 --warmup 1 --eager --no-extras --no-cpu-baseline > gpurun_out/final_launches.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -c 45 -o gpurun_out/final_step_full -f python bench.py --steps 2 --warmup 1 --eager --no-extras --no-cpu-baseline > gpurun_out/final_ncu.log 2>&1
python - <<PY
import json
for f in ("final_bench", "final_bench_ref", "final_bench_pdl"):
    try:
        d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
        print(f, d.get("ms_per_step"), d.get("value"), (d.get("e2e") or {}).get("value"))
    except Exception as e:
        print(f, "failed", e)
PY
