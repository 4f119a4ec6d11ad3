#!/bin/bash
# N-GPU A/B of the early exchange: reduce-scatter (default for world > 2) vs all-to-all push kernel vs copy engines
cd "$(dirname "$0")/../.." && mkdir -p gpurun_out && O=gpurun_out
G=${1:-4}
for cfg in "sm 1" "sm 0" "dma 1"; do set -- $cfg
T4K_DP_EARLY=$1 T4K_DP_RS=$2 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $G --steps 100 --no-extras --no-cpu-baseline 2>$O/r02_ab.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1 rs=$2', d['n_gpus'], round(d['ms_per_step']*1e3,2), 'us e2e', round(d['e2e']['value']/1e6,3), 'loss', d['final_loss'], (d.get('extras') or {}).get('mnist_strong_scaling',{}).get('ms_per_step'))"
done
