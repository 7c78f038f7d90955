#!/bin/bash
# water1m step time under a list of environment settings at N GPUs.  usage: bash tools/gpu_sweep_big.sh <tag> <N> "A=1" "A=2" ...
tag=$1; N=$2; shift; shift
mkdir -p gpurun_out
i=0
for cfg in "$@"; do
  i=$((i+1))
  if [ "$N" = 1 ]; then
    env $cfg timeout 300 python bench.py --workload water1m --steps 5 --warmup 3 > gpurun_out/${tag}_$i.json 2> gpurun_out/${tag}_$i.err
  else
    env $cfg timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29800+i)) bench.py --gpus $N --workload water1m --steps 5 --warmup 3 > gpurun_out/${tag}_$i.json 2> gpurun_out/${tag}_$i.err
  fi
  echo "N=$N [$cfg] $(python -c "import json; d=json.loads(open('gpurun_out/${tag}_$i.json').read().strip().splitlines()[-1]); print('ms/step', round(d['ms_per_step'],3), 'induce', round(d['ms_per_induce'],3), 'e2e', round(d['e2e']['ms_per_step'],3))")"
done
