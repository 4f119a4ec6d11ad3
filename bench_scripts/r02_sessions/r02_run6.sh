#!/bin/bash
cd "$(dirname "$0")/../.." && mkdir -p gpurun_out && O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r02_t9.log 2>&1; echo "rc=$?" >> $O/r02_t9.log; tail -15 $O/r02_t9.log
timeout 900 python bench.py --steps 200 --no-cpu-baseline > $O/r02_bench5.json 2> $O/r02_bench5.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('$O/r02_bench5.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['final_loss']); [print(c) for c in d['calls']]; print(d.get('extras',{}).get('gan_t4_40b'))"
