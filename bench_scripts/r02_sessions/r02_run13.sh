#!/bin/bash
cd "$(dirname "$0")/../.." && mkdir -p gpurun_out && O=gpurun_out
for m in dma sm; do
T4K_DP_EARLY=$m timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 100 --no-extras --no-cpu-baseline > $O/r02_b8_$m.json 2> $O/r02_b8_$m.err
done
python -c "
import json
for m in ('dma','sm'):
    d=json.load(open('$O/r02_b8_%s.json' % m)); print(m, d['n_gpus'], round(d['ms_per_step']*1e3,2),'us', round(d['value']/1e6,3),'M/s e2e', round(d['e2e']['value']/1e6,3), d['config']['exchange'], d['timing']['window_ms'][:3], d['launches_per_step'], d['final_loss'])
"
