#!/bin/bash
# dhfr2 step time under a list of environment settings (one GPU).  usage: bash tools/gpu_sweep.sh <tag> "A=1 B=2" "A=2" ...
tag=$1; shift
mkdir -p gpurun_out
i=0
for cfg in "$@"; do
  i=$((i+1))
  env $cfg timeout 200 python bench.py --steps 40 --warmup 8 --no-cpu > gpurun_out/${tag}_$i.json 2> gpurun_out/${tag}_$i.err
  echo "[$cfg] $(python -c "import json; d=json.loads(open('gpurun_out/${tag}_$i.json').read().strip().splitlines()[-1]); print('ms/step', round(d['ms_per_step'],4), 'induce', round(d['ms_per_induce'],4), 'e2e', round(d['e2e']['ms_per_step'],4))")"
done
