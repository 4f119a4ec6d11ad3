#!/bin/bash
# round-2 evidence run (1 GPU): full GPU test suite, bench lines (ours + reference arm), ncu launch list of the captured step, ncu --set full of
# the step's dominant kernels and of the 4096^3 GEMM.  Everything lands in gpurun_out/ (< 64 MiB); summaries are copied into profiles/ afterwards.
cd "$(dirname "$0")/.." && mkdir -p gpurun_out && O=gpurun_out
timeout 1200 python -m pytest tests -q -m gpu > $O/r02_final_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r02_final_pytest.log
timeout 600 python bench.py > $O/r02_final_bench.json 2> $O/r02_final_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --steps 20 --warmup 3 > $O/r02_final_bench_s20.json 2> $O/r02_final_bench_s20.err; echo "bench20 rc=$?"
timeout 300 python bench.py --impl reference > $O/r02_final_bench_ref.json 2> $O/r02_final_bench_ref.err; echo "ref rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_final_launches.csv python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > $O/r02_final_launches.log 2>&1
python tests/parse_launches.py $O/r02_final_launches.csv | tail -12
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_gemm_tl|k_cpr2_bwd|k_cpr2_fwd" -s 24 -c 4 -o $O/r02_final_step -f python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > $O/r02_final_ncu_step.log 2>&1; tail -2 $O/r02_final_ncu_step.log | cut -c1-200
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_gemm_tc2|k_pack" -c 6 -o $O/r02_final_gemm4096 -f python bench_scripts/gemm4096_once.py > $O/r02_final_ncu_gemm.log 2>&1; tail -2 $O/r02_final_ncu_gemm.log
ls -la $O/r02_final_*
