#!/bin/bash
cd "$(dirname "$0")/../.." && mkdir -p gpurun_out && O=gpurun_out
G=${1:-8}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $G --steps 100 --no-extras --no-cpu-baseline 2>$O/r02_ab.err > $O/r02_scale_${G}_final.json; python -c "import json,sys; d=json.load(open('$O/r02_scale_${G}_final.json')); print(d['n_gpus'], round(d['ms_per_step']*1e3,2), 'us e2e', round(d['e2e']['value']/1e6,3), 'loss', d['final_loss'], (d.get('extras') or {}).get('mnist_strong_scaling',{}).get('ms_per_step'))"
