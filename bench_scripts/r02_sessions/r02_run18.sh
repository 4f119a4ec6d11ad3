#!/bin/bash
cd "$(dirname "$0")/../.." && mkdir -p gpurun_out && O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_dp_lanes.py tests/test_gpu_dp_multi.py -m gpu -q 2>&1 | tail -3
for cfg in "sm 1" "sm 0" "dma 1"; do set -- $cfg
T4K_DP_EARLY=$1 T4K_DP_REST=$2 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 100 --no-extras --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1 rest=$2', d['n_gpus'], round(d['ms_per_step']*1e3,2), round(d['e2e']['value']/1e6,3), d['final_loss'])"
done
