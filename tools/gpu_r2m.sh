#!/bin/bash
# round 2, multi-GPU visit (gpurun --gpus N): NCCL parity check of the decomposed path, decomposed 1 M-atom bench, operator CTA cap
# A/B, rank-0 timeline.  usage: bash tools/gpu_r2m.sh <N> [full]
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
nvidia-smi -L | head -8
timeout 300 $TR --master-port 29611 tools/nccl_check.py water30 > gpurun_out/r02m_n${N}_nccl_check.log 2>&1
echo "nccl_check water30 rc=$?"; grep RESULT gpurun_out/r02m_n${N}_nccl_check.log; tail -2 gpurun_out/r02m_n${N}_nccl_check.log
timeout 400 $TR --master-port 29613 bench.py --gpus $N --workload water1m --steps 5 --warmup 3 --no-cpu > gpurun_out/r02m_water1m_n$N.json 2> gpurun_out/r02m_water1m_n$N.err
echo "water1m N=$N rc=$?"; tail -2 gpurun_out/r02m_water1m_n$N.err
if [ -n "$2" ]; then
  APX_TL_CTAS=6 timeout 400 $TR --master-port 29614 bench.py --gpus $N --workload water1m --steps 5 --warmup 3 --no-cpu > gpurun_out/r02m_water1m_n${N}_ct6.json 2> gpurun_out/r02m_water1m_n${N}_ct6.err
  APX_DIST_P2P=0 timeout 400 $TR --master-port 29615 bench.py --gpus $N --workload water1m --steps 5 --warmup 3 --no-cpu > gpurun_out/r02m_water1m_n${N}_nccl.json 2> gpurun_out/r02m_water1m_n${N}_nccl.err
  timeout 300 $TR --master-port 29616 tools/trace_step.py --workload water1m --steps 2 --out gpurun_out/r02m_trace_water1m_n$N.txt > gpurun_out/r02m_trace_n$N.log 2>&1
  head -40 gpurun_out/r02m_trace_water1m_n$N.txt
fi
for f in gpurun_out/r02m_water1m_n${N}*.json; do
  python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "n_gpus", d["n_gpus"], "ms/step", round(d["ms_per_step"],3), "induce", round(d["ms_per_induce"],3), "iters", d["pcg_iterations"], "cfg", d["config"].get("parallelism"), json.dumps(d.get("decomposition", d.get("dist", "")))[:600])
except Exception as e:
    print(sys.argv[1], "ERR", e)
PY
done
