#!/bin/bash
# N-GPU scaling of the headline workload on one box (quick flags: no CPU baseline / DPP / secondary configs)
# usage: tools/scale_run.sh [N ...]   (default 1 2 4 8)
F="--steps 2 --warmup 1 --no-cpu-baseline --no-dpp --no-configs --no-incremental-extra"
for n in ${@:-1 2 4 8}; do
  if [ $n -eq 1 ]; then python bench.py --gpus 1 $F > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n $F > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err; fi
  python -c "
import json
d=json.loads(open('gpurun_out/scale_n$n.json').read().strip().splitlines()[-1])
print('N=$n', round(d['value']), 'BO it/s', d['ms_per_step'], 'ms/step', d['clocks'])"
done
nvidia-smi topo -m | head -12
