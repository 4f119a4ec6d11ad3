#!/bin/bash
# round 2, GPU call 1: first contact of the layer GEMM + the evidence the round-1 verdict asked for (GEMM 4096^3 ncu, GAN launch list)
cd "$(dirname "$0")/../.." && mkdir -p gpurun_out && O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/r02_run1_smi.txt 2>&1
timeout 180 python bench_scripts/tl_smoke.py > $O/r02_tl_smoke.txt 2>&1; rc=$?; echo "tl_smoke rc=$rc" | tee -a $O/r02_tl_smoke.txt
if [ $rc -ne 0 ]; then export T4K_GEMM_TL=0; echo "TL engine disabled for the rest of this run"; fi
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -k "gemm or linear or mlp_head" > $O/r02_t1.log 2>&1; echo "rc=$?" >> $O/r02_t1.log; tail -5 $O/r02_t1.log
timeout 300 python bench_scripts/gemm_probe.py > $O/r02_gemm_probe.txt 2>&1
T4K_TCF_CLUSTER=1 timeout 300 python bench_scripts/gemm_probe.py > $O/r02_gemm_probe_tcfcluster.txt 2>&1
timeout 1200 python -m pytest tests/test_gpu_model.py -x -q > $O/r02_t2.log 2>&1; echo "rc=$?" >> $O/r02_t2.log; tail -5 $O/r02_t2.log
timeout 900 python bench.py --steps 200 > $O/r02_bench1.json 2> $O/r02_bench1.err; echo "bench rc=$?"
timeout 600 python bench.py --steps 20 --no-extras --no-cpu-baseline > $O/r02_bench1_s20.json 2>> $O/r02_bench1.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/r02_launches_gan.csv python bench_scripts/gan_step.py 3 > $O/r02_launches_gan.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_launches_step.csv python bench.py --steps 2 --warmup 3 --eager --no-extras --no-cpu-baseline > $O/r02_launches_step.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_gemm_tc|k_pack" -c 8 -o $O/r02_gemm4096 python bench_scripts/gemm4096_once.py > $O/r02_ncu_gemm4096.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_gemm_tl" -c 6 -o $O/r02_gemm_tl python bench.py --steps 2 --warmup 3 --eager --no-extras --no-cpu-baseline > $O/r02_ncu_gemm_tl.log 2>&1
ls -la $O | tail -30
