#!/bin/bash
cd "$(dirname "$0")/../.." && mkdir -p gpurun_out && O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_dp_multi.py tests/test_gpu_model.py -m gpu -q -k "two_gpu or optimizer_state or dcgan" 2>&1 | grep -v "^$" > $O/r02_t16.log; grep -n "^E  .*Error\|^E   .*assert\|^FAILED\|passed\|failed" $O/r02_t16.log | head -20
timeout 300 python bench.py --steps 100 --no-extras --no-cpu-baseline > $O/r02_b1.json 2> $O/r02_b1.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 100 --no-extras --no-cpu-baseline > $O/r02_b2.json 2> $O/r02_b2.err
python -c "
import json
for f in ('$O/r02_b1.json','$O/r02_b2.json'):
    d=json.load(open(f)); print(d['n_gpus'], round(d['ms_per_step']*1e3,2),'us', round(d['value']/1e6,3),'M/s e2e', round(d['e2e']['value']/1e6,3), d['config']['exchange'], d['timing']['window_ms'][:3])
"
