#!/bin/bash
cd "$(dirname "$0")/../.." && mkdir -p gpurun_out && O=gpurun_out
timeout 120 python bench_scripts/tl_trace.py > $O/r02_tl_trace.txt 2>&1; cat $O/r02_tl_trace.txt
timeout 1500 python -m pytest tests/test_gpu_kernels.py -x -q > $O/r02_t3.log 2>&1; echo "rc=$?" >> $O/r02_t3.log; tail -4 $O/r02_t3.log
timeout 1200 python -m pytest tests/test_gpu_model.py -x -q > $O/r02_t4.log 2>&1; echo "rc=$?" >> $O/r02_t4.log; tail -4 $O/r02_t4.log
timeout 300 python bench_scripts/gemm_probe.py > $O/r02_gemm_probe2.txt 2>&1; head -20 $O/r02_gemm_probe2.txt
timeout 900 python bench.py --steps 200 > $O/r02_bench2.json 2> $O/r02_bench2.err; echo "bench rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_launches_step2.csv python bench.py --steps 2 --warmup 3 --eager --no-extras --no-cpu-baseline > $O/r02_launches_step2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_gemm_tl" -c 6 -o $O/r02_gemm_tl python bench.py --steps 2 --warmup 3 --eager --no-extras --no-cpu-baseline > $O/r02_ncu_gemm_tl.log 2>&1
