#!/bin/bash
mkdir -p gpurun_out
timeout 170 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench rc=$?"
timeout 60 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 2 --warmup 1 --eager --no-extras --no-cpu-baseline > gpurun_out/final_launches.log 2>&1; echo "launches rc=$?"
timeout 60 ncu --set full --clock-control none -k regex:k_cpr2_bwd -s 1 -c 1 -o gpurun_out/final_cpr2_bwd -f python bench.py --steps 2 --warmup 1 --eager --no-extras --no-cpu-baseline > gpurun_out/final_ncu.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out | tail -8
