#!/bin/bash
# round-end evidence run (1 GPU): full GPU test suite, bench lines (ours + reference arm), ncu launch list of one eager step and
# ncu --set full of the dominant kernel.  Keep what lands in gpurun_out/ under 64 MiB (gpurun drops the whole directory above that:
# a --set full capture of every launch of the step is ~65 MB, hence the -k filter).
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/final_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/final_pytest.log
timeout 300 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference > gpurun_out/final_bench_ref.json 2> gpurun_out/final_bench_ref.err; echo "ref rc=$?"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 2 --warmup 1 --eager --no-extras --no-cpu-baseline > gpurun_out/final_launches.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_cpr2_bwd -s 1 -c 1 -o gpurun_out/final_cpr2_bwd -f python bench.py --steps 2 --warmup 1 --eager --no-extras --no-cpu-baseline > gpurun_out/final_ncu.log 2>&1
python tests/parse_launches.py gpurun_out/final_launches.csv | tail -14
