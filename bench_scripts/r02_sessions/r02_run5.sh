#!/bin/bash
cd "$(dirname "$0")/../.." && mkdir -p gpurun_out && O=gpurun_out
timeout 200 python bench_scripts/tl_smoke.py > $O/r02_tl_smoke3.txt 2>&1; tail -2 $O/r02_tl_smoke3.txt
timeout 1500 python -m pytest tests/test_gpu_kernels.py -x -q -k "gemm or linear or head or optim or conv_pool or step_graph" > $O/r02_t7.log 2>&1; echo "rc=$?" >> $O/r02_t7.log; tail -4 $O/r02_t7.log
timeout 1200 python -m pytest tests/test_gpu_model.py -x -q > $O/r02_t8.log 2>&1; echo "rc=$?" >> $O/r02_t8.log; tail -4 $O/r02_t8.log
timeout 120 python bench_scripts/tl_trace.py > $O/r02_tl_trace3.txt 2>&1; grep -A1 "L2" $O/r02_tl_trace3.txt | head -20
timeout 300 python bench_scripts/gemm_probe.py > $O/r02_gemm_probe4.txt 2>&1; head -12 $O/r02_gemm_probe4.txt | cut -c1-40,150-260
timeout 900 python bench.py --steps 200 --no-cpu-baseline > $O/r02_bench4.json 2> $O/r02_bench4.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('$O/r02_bench4.json')); print(d['value'], d['ms_per_step'], d['e2e']['value']); [print(c) for c in d['calls']]; print(d.get('extras',{}).get('gan_t4_40b'))"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_launches_step4.csv python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > $O/r02_launches_step4.log 2>&1
python tests/parse_launches.py $O/r02_launches_step4.csv
