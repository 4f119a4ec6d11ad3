#!/bin/bash
# weak-scaling lines at 1/2/4/8 GPUs of one box (what the driver's SCALE step runs), steps 100
cd "$(dirname "$0")/.." && mkdir -p gpurun_out && O=gpurun_out
timeout 300 python bench.py --gpus 1 --steps 100 --no-extras --no-cpu-baseline > $O/r02_scale_1.json 2> $O/r02_scale_1.err
for g in 2 4 8; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $g --steps 100 --no-extras --no-cpu-baseline > $O/r02_scale_$g.json 2> $O/r02_scale_$g.err
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 20 > $O/r02_scale_8_full.json 2> $O/r02_scale_8_full.err
python -c "
import json
b=None
for g in (1,2,4,8):
    d=json.load(open('$O/r02_scale_%d.json' % g)); b=b or d['value']
    print(d['n_gpus'], round(d['ms_per_step']*1e3,2),'us', round(d['value']/1e6,3),'M/s eff', round(d['value']/b/g,3), 'e2e', round(d['e2e']['value']/1e6,3), d['config']['exchange'], d['launches_per_step'], d['final_loss'])
d=json.load(open('$O/r02_scale_8_full.json')); print('8 full', round(d['ms_per_step']*1e3,2), d.get('extras',{}).get('gan_t4_40b'), d.get('extras',{}).get('conv2d_3x3_64',{}).get('fwd'))
"
