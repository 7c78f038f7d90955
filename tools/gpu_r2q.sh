#!/bin/bash
# round 2, multi-GPU visit q (gpurun --gpus N): direct transport (APX_DIST_P2P=3) over real NVLink: parity, 1 M-atom bench against
# the windowed transport (2), CTA sweep of the exchange kernel, rank-0 timeline.  usage: bash tools/gpu_r2q.sh <N> [full]
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
nvidia-smi -L | head -8
timeout 300 $TR --master-port 29611 tools/nccl_check.py water30 > gpurun_out/r02q_n${N}_check.log 2>&1
echo "nccl_check(p2p=3) water30 rc=$?"; grep RESULT gpurun_out/r02q_n${N}_check.log; tail -2 gpurun_out/r02q_n${N}_check.log
timeout 300 python tools/direct_check.py --world $N --blob water30 --rep 2x2x2 --timeout 250 > gpurun_out/r02q_n${N}_direct.log 2>&1
echo "direct_check rc=$?"; grep RESULT gpurun_out/r02q_n${N}_direct.log
run() { # tag, env...
  tag=$1; shift
  env "$@" timeout 400 $TR --master-port 29613 bench.py --gpus $N --workload water1m --steps 5 --warmup 3 --no-cpu > gpurun_out/r02q_water1m_n${N}_$tag.json 2> gpurun_out/r02q_water1m_n${N}_$tag.err
  echo "water1m N=$N $tag rc=$?"; tail -1 gpurun_out/r02q_water1m_n${N}_$tag.err
}
run p3 APX_DIST_P2P=3
if [ -n "$2" ]; then
  run p2 APX_DIST_P2P=2
  run p3_c296 APX_DIST_P2P=3 APX_DX_CTAS=296
  run p3_c1184 APX_DIST_P2P=3 APX_DX_CTAS=1184
  timeout 300 $TR --master-port 29616 tools/trace_step.py --workload water1m --steps 2 --out gpurun_out/r02q_trace_water1m_n$N.txt > gpurun_out/r02q_trace_n$N.log 2>&1
  head -45 gpurun_out/r02q_trace_water1m_n$N.txt
fi
for f in gpurun_out/r02q_water1m_n${N}_*.json; do
  python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "n_gpus", d["n_gpus"], "ms/step", round(d["ms_per_step"],3), "induce", round(d["ms_per_induce"],3), "iters", d["pcg_iterations"], json.dumps(d.get("decomposition", d.get("dist", "")))[:500])
except Exception as e:
    print(sys.argv[1], "ERR", e)
PY
done
