#!/bin/bash
cd "$(dirname "$0")/../.." && mkdir -p gpurun_out && O=gpurun_out
for z in 0 1 0 1; do
T4K_CPR_ZERO=$z timeout 600 python bench.py --steps 100 --no-cpu-baseline --no-extras > $O/r02_b1_z$z.json 2> $O/r02_b1.err
python -c "
import json
d=json.load(open('$O/r02_b1_z$z.json')); print('zero_all=$z', round(d['ms_per_step']*1e3,2),'us e2e', round(d['e2e']['value']/1e6,3), [c['us'] for c in d['calls']])
"
done
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -m gpu -q -x -k "conv_pool or fused_block or model_train or step_graph" 2>&1 | tail -2
