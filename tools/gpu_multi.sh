#!/bin/bash
# Multi-GPU visit (gpurun --gpus N): NCCL parity check, decomposed 1 M-atom bench at N and at 1, replica bench of dhfr2.
# usage: bash tools/gpu_multi.sh <tag> <N> [pytest-args]
tag=${1:-multi}; N=${2:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
if [ -n "$3" ]; then
  timeout 900 python -m pytest $3 -m gpu -q -x > gpurun_out/${tag}_pytest.log 2>&1
  echo "pytest rc=$?"; tail -4 gpurun_out/${tag}_pytest.log
fi
timeout 300 $TR --master-port 29611 tools/nccl_check.py water30 > gpurun_out/${tag}_nccl_check.log 2>&1
echo "nccl_check water30 rc=$?"; grep RESULT gpurun_out/${tag}_nccl_check.log; tail -3 gpurun_out/${tag}_nccl_check.log
timeout 300 $TR --master-port 29612 tools/nccl_check.py dhfr2 > gpurun_out/${tag}_nccl_check_dhfr2.log 2>&1
echo "nccl_check dhfr2 rc=$?"; grep RESULT gpurun_out/${tag}_nccl_check_dhfr2.log; tail -3 gpurun_out/${tag}_nccl_check_dhfr2.log
for wl in water1m water96k; do
  timeout 400 $TR --master-port 29613 bench.py --gpus $N --workload $wl --steps 5 --warmup 3 > gpurun_out/${tag}_${wl}_n$N.json 2> gpurun_out/${tag}_${wl}_n$N.err
  echo "$wl N=$N rc=$?"; cat gpurun_out/${tag}_${wl}_n$N.json; tail -3 gpurun_out/${tag}_${wl}_n$N.err
done
timeout 300 python bench.py --workload water1m --steps 5 --warmup 3 > gpurun_out/${tag}_water1m_n1.json 2> gpurun_out/${tag}_water1m_n1.err
echo "water1m N=1 rc=$?"; cat gpurun_out/${tag}_water1m_n1.json
timeout 300 $TR --master-port 29614 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu > gpurun_out/${tag}_dhfr2_n$N.json 2> gpurun_out/${tag}_dhfr2_n$N.err
echo "dhfr2 replicas N=$N rc=$?"; cat gpurun_out/${tag}_dhfr2_n$N.json; tail -3 gpurun_out/${tag}_dhfr2_n$N.err
