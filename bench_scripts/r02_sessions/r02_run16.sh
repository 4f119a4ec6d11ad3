#!/bin/bash
cd "$(dirname "$0")/../.." && mkdir -p gpurun_out && O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -m gpu -q -x -k "conv_pool or fused_block or model_train or step_graph or dataset or dropout or gan" 2>&1 | grep -v "^$" > $O/r02_t21.log; grep -n "^E  .*Error\|^E   .*assert\|^FAILED\|passed\|failed" $O/r02_t21.log | head -20
timeout 600 python bench.py --steps 100 --no-cpu-baseline > $O/r02_b1_h.json 2> $O/r02_b1.err
python -c "
import json
d=json.load(open('$O/r02_b1_h.json')); g=d['extras']['gan_t4_40b']; print(round(d['ms_per_step']*1e3,2),'us e2e', round(d['e2e']['value']/1e6,3), '| GAN', g['ms_per_iteration'], 'ms', g['launches_per_iteration'], 'launches'); [print(c['call'][:40], c['us']) for c in d['calls']]
"
