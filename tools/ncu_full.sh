#!/bin/bash
# ncu --set full capture of selected kernels of one short bench run (one GPU).
# usage (under gpurun): bash tools/ncu_full.sh <tag> <kernel-regex> [skip] [count]
tag=$1; rx=$2; skip=${3:-10}; cnt=${4:-6}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$rx" -s $skip -c $cnt -f -o gpurun_out/${tag} \
   python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/${tag}_ncu_full.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/${tag}_ncu_full.log
