#!/bin/bash
cd "$(dirname "$0")/../.." && mkdir -p gpurun_out && O=gpurun_out
for z in 0 1 0 1; do
T4K_OPT_LATE=$z timeout 300 python bench.py --steps 100 --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('opt_late=$z', round(d['ms_per_step']*1e3,2),'us e2e', round(d['e2e']['value']/1e6,3), d['final_loss'], d['launches_per_step'])"
done
T4K_OPT_LATE=1 timeout 300 python -m pytest tests/test_gpu_model.py -m gpu -q -x -k "step_graph or dataset" 2>&1 | tail -2
