#!/bin/bash
cd "$(dirname "$0")/../.." && mkdir -p gpurun_out && O=gpurun_out
for mi in 1 0; do
T4K_DP_MIRROR=$mi timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 100 --no-extras --no-cpu-baseline 2>$O/r02_ab.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('mirror=$mi', d['n_gpus'], round(d['ms_per_step']*1e3,2), 'us e2e', round(d['e2e']['value']/1e6,3), 'M/s =', round(512*2/d['e2e']['value']*1e6,2), 'us/step; final_loss', d['final_loss'], 'host-read loss', d['e2e']['last_loss_read_on_host'])"
done
