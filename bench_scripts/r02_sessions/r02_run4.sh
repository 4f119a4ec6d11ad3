#!/bin/bash
cd "$(dirname "$0")/../.." && mkdir -p gpurun_out && O=gpurun_out
timeout 120 python bench_scripts/tl_trace.py > $O/r02_tl_trace2.txt 2>&1; cat $O/r02_tl_trace2.txt
timeout 1500 python -m pytest tests/test_gpu_kernels.py -x -q > $O/r02_t5.log 2>&1; echo "rc=$?" >> $O/r02_t5.log; tail -4 $O/r02_t5.log
timeout 1200 python -m pytest tests/test_gpu_model.py -x -q > $O/r02_t6.log 2>&1; echo "rc=$?" >> $O/r02_t6.log; tail -4 $O/r02_t6.log
timeout 300 python bench_scripts/gemm_probe.py > $O/r02_gemm_probe3.txt 2>&1; head -12 $O/r02_gemm_probe3.txt | cut -c1-40,150-260
timeout 900 python bench.py --steps 200 > $O/r02_bench3.json 2> $O/r02_bench3.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('$O/r02_bench3.json')); print(d['value'], d['ms_per_step'], d['e2e']['value']); [print(c) for c in d['calls']]; print(d.get('extras',{}).get('gan_t4_40b'))"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_launches_step3.csv python bench.py --steps 2 --warmup 3 --eager --no-extras --no-cpu-baseline > $O/r02_launches_step3.log 2>&1
python tests/parse_launches.py $O/r02_launches_step3.csv
