#!/bin/bash
# round 2, visit r (gpurun --gpus 8): direct transport at 8 GPUs: parity on the 2x2x2 water box, decomposed 1 M-atom bench, rank-0 timeline
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
nvidia-smi -L | head -8
nvidia-smi topo -m 2>/dev/null | head -12
timeout 300 $TR --master-port 29611 tools/nccl_check.py water30 2x2x2 > gpurun_out/r02r_n${N}_check.log 2>&1
echo "nccl_check(p2p=3) water30 2x2x2 rc=$?"; grep RESULT gpurun_out/r02r_n${N}_check.log | cut -c1-160; tail -2 gpurun_out/r02r_n${N}_check.log
run() { # tag, env...
  tag=$1; shift
  env "$@" timeout 400 $TR --master-port 29613 bench.py --gpus $N --workload water1m --steps 5 --warmup 3 --no-cpu > gpurun_out/r02r_water1m_n${N}_$tag.json 2> gpurun_out/r02r_water1m_n${N}_$tag.err
  echo "water1m N=$N $tag rc=$?"; tail -1 gpurun_out/r02r_water1m_n${N}_$tag.err
}
run p3 APX_DIST_P2P=3
run p3_c148 APX_DIST_P2P=3 APX_DX_CTAS=148
timeout 300 $TR --master-port 29616 tools/trace_step.py --workload water1m --steps 2 --out gpurun_out/r02r_trace_water1m_n$N.txt > gpurun_out/r02r_trace_n$N.log 2>&1
head -45 gpurun_out/r02r_trace_water1m_n$N.txt
for f in gpurun_out/r02r_water1m_n${N}_*.json; do
  python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "n_gpus", d["n_gpus"], "ms/step", round(d["ms_per_step"],3), "induce", round(d["ms_per_induce"],3), "iters", d["pcg_iterations"])
except Exception as e:
    print(sys.argv[1], "ERR", e)
PY
done
