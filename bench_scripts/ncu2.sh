#!/bin/bash
mkdir -p gpurun_out
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_cpr2_fwd -s 3 -c 1 -o gpurun_out/ncu_cpr2_fwd -f python tests/perf_probe.py cpr > gpurun_out/ncu_a.log 2>&1; tail -1 gpurun_out/ncu_a.log
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_head_bwd -s 3 -c 1 -o gpurun_out/ncu_head_bwd -f python tests/perf_probe.py fc > gpurun_out/ncu_b.log 2>&1; tail -1 gpurun_out/ncu_b.log
