#!/bin/bash
# fused push/pull windows (APX_DIST_P2P=2) against copy-engine windows (1) and NCCL (0): parity, then the 1 M-atom bench.
tag=${1:-p2p2}; N=${2:-2}; modes=${3:-"2 1 0"}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
APX_DIST_P2P=2 timeout 240 $TR --master-port 29911 tools/nccl_check.py water30 > gpurun_out/${tag}_check_water30.log 2>&1
echo "check water30 rc=$?"; grep RESULT gpurun_out/${tag}_check_water30.log | sort -u | cut -c1-150; tail -2 gpurun_out/${tag}_check_water30.log | cut -c1-300
APX_DIST_P2P=2 timeout 240 $TR --master-port 29912 tools/nccl_check.py dhfr2 > gpurun_out/${tag}_check_dhfr2.log 2>&1
echo "check dhfr2 rc=$?"; grep RESULT gpurun_out/${tag}_check_dhfr2.log | sort -u | cut -c1-150; tail -2 gpurun_out/${tag}_check_dhfr2.log | cut -c1-300
for v in $modes; do
  APX_DIST_P2P=$v timeout 300 $TR --master-port $((29913+v)) bench.py --gpus $N --workload water1m --steps 5 --warmup 3 > gpurun_out/${tag}_water1m_n${N}_p2p$v.json 2> gpurun_out/${tag}_water1m_n${N}_p2p$v.err
  echo "P2P=$v water1m N=$N rc=$? $(python -c "import json; d=json.loads(open('gpurun_out/${tag}_water1m_n${N}_p2p$v.json').read().strip().splitlines()[-1]); print('ms/step', round(d['ms_per_step'],3), 'induce', round(d['ms_per_induce'],3), 'e2e', round(d['e2e']['ms_per_step'],3))")"
done
