#!/bin/bash
# N-GPU decomposed 1 M-atom bench under a few NCCL point-to-point channel settings.
# usage: bash tools/gpu_nccl_tune.sh <tag> <N>
tag=${1:-tune}; N=${2:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
i=0
for cfg in "" "NCCL_MIN_P2P_NCHANNELS=8" "NCCL_MIN_P2P_NCHANNELS=16" "NCCL_MIN_P2P_NCHANNELS=32 NCCL_MAX_P2P_NCHANNELS=32" "NCCL_MIN_P2P_NCHANNELS=16 NCCL_BUFFSIZE=16777216"; do
  i=$((i+1))
  env $cfg timeout 300 $TR --master-port $((29630+i)) bench.py --gpus $N --workload water1m --steps 5 --warmup 3 > gpurun_out/${tag}_cfg$i.json 2> gpurun_out/${tag}_cfg$i.err
  echo "cfg$i [$cfg] rc=$? $(python -c "import json,sys; d=json.loads(open('gpurun_out/${tag}_cfg$i.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['ms_per_induce'])")"
done
