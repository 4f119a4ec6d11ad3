#!/bin/bash
cd "$(dirname "$0")/../.." && mkdir -p gpurun_out && O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -m gpu -q -x -k "linear or gan or model_train or step_graph" 2>&1 | grep -v "^$" > $O/r02_t20.log; grep -n "^E  .*Error\|^E   .*assert\|^FAILED\|passed\|failed" $O/r02_t20.log | head -20
for pr in 1 0; do
T4K_TL_PAIR=$pr timeout 600 python bench.py --steps 100 --no-cpu-baseline > $O/r02_b1_pair$pr.json 2> $O/r02_b1.err
done
python -c "
import json
for pr in (1,0):
    d=json.load(open('$O/r02_b1_pair%d.json' % pr)); g=d['extras']['gan_t4_40b']; print('pair',pr, round(d['ms_per_step']*1e3,2),'us | GAN', g['ms_per_iteration'], 'ms', g['launches_per_iteration'], 'launches eager', g['ms_per_iteration_eager'], g['losses_after'])
"
