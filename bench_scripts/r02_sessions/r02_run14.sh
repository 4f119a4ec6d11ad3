#!/bin/bash
cd "$(dirname "$0")/../.." && mkdir -p gpurun_out && O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_dp_lanes.py tests/test_gpu_dp_multi.py -m gpu -q 2>&1 | grep -v "^$" > $O/r02_t19.log; grep -n "^E  .*Error\|^E   .*assert\|^FAILED\|passed\|failed" $O/r02_t19.log | head -20
for g in 2 4 8; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $g --steps 100 --no-extras --no-cpu-baseline > $O/r02_bN$g.json 2> $O/r02_bN$g.err
done
python -c "
import json
for g in (2,4,8):
    d=json.load(open('$O/r02_bN%d.json' % g)); print(d['n_gpus'], round(d['ms_per_step']*1e3,2),'us', round(d['value']/1e6,3),'M/s e2e', round(d['e2e']['value']/1e6,3), d['config']['exchange'], d['timing']['window_ms'][:3], d['launches_per_step'], d['final_loss'])
"
