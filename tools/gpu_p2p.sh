#!/bin/bash
# N-GPU check of the peer-memory transport: parity (tools/nccl_check.py), then the decomposed 1 M-atom bench with and without it.
tag=${1:-p2p}; N=${2:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 240 $TR --master-port 29711 tools/nccl_check.py water30 > gpurun_out/${tag}_check_water30.log 2>&1
echo "check water30 rc=$?"; grep RESULT gpurun_out/${tag}_check_water30.log | sort -u | cut -c1-200; tail -2 gpurun_out/${tag}_check_water30.log | cut -c1-300
timeout 240 $TR --master-port 29712 tools/nccl_check.py dhfr2 > gpurun_out/${tag}_check_dhfr2.log 2>&1
echo "check dhfr2 rc=$?"; grep RESULT gpurun_out/${tag}_check_dhfr2.log | sort -u | cut -c1-200; tail -2 gpurun_out/${tag}_check_dhfr2.log | cut -c1-300
for v in 1 0; do
  APX_DIST_P2P=$v timeout 300 $TR --master-port $((29713+v)) bench.py --gpus $N --workload water1m --steps 5 --warmup 3 > gpurun_out/${tag}_water1m_n${N}_p2p$v.json 2> gpurun_out/${tag}_water1m_n${N}_p2p$v.err
  echo "P2P=$v water1m N=$N rc=$? $(python -c "import json; d=json.loads(open('gpurun_out/${tag}_water1m_n${N}_p2p$v.json').read().strip().splitlines()[-1]); print('ms/step', round(d['ms_per_step'],3), 'induce', round(d['ms_per_induce'],3), 'e2e', round(d['e2e']['ms_per_step'],3))")"
  tail -2 gpurun_out/${tag}_water1m_n${N}_p2p$v.err | cut -c1-300
done
