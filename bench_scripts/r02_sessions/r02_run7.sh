#!/bin/bash
cd "$(dirname "$0")/../.." && mkdir -p gpurun_out && O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_comm.py tests/test_gpu_dp_lanes.py -m gpu -q -x 2>&1 | grep -v "^$" > $O/r02_t12.log; grep -n "^E  .*Error\|^E   \|^FAILED\|passed\|failed" $O/r02_t12.log | head -30
timeout 900 python -m pytest tests/test_gpu_model.py -m gpu -q 2>&1 | grep -v "^$" > $O/r02_t13.log; grep -n "^E  .*Error\|^FAILED\|passed\|failed" $O/r02_t13.log | head -30
timeout 900 python bench.py --steps 20 --no-cpu-baseline > $O/r02_bench6.json 2> $O/r02_bench6.err; echo "bench rc=$?"; tail -3 $O/r02_bench6.err; python -c "
import json; d=json.load(open('$O/r02_bench6.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['final_loss'], d['timing'], d['gpu_launches']); [print(c) for c in d['calls']]; print(d.get('extras',{}).get('gan_t4_40b'))"
