#!/bin/bash
# round 2, visit y (gpurun --gpus 8): final 8-GPU check of the direct transport: parity, bulk-engine against copy-loop exchange kernel
N=8
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29611 tools/nccl_check.py water30 2x2x2 > gpurun_out/r02y_n${N}_check.log 2>&1
echo "nccl_check(p2p=3, bulk) rc=$?"; grep RESULT gpurun_out/r02y_n${N}_check.log | cut -c1-150 | head -3; tail -2 gpurun_out/r02y_n${N}_check.log | cut -c1-300
run() { # tag, env...
  tag=$1; shift
  env "$@" timeout 400 $TR --master-port 29613 bench.py --gpus $N --workload water1m --steps 5 --warmup 3 --no-cpu > gpurun_out/r02y_water1m_n${N}_$tag.json 2> gpurun_out/r02y_water1m_n${N}_$tag.err
  echo "water1m N=$N $tag rc=$?"; grep "apx dx trace\] rank 0" gpurun_out/r02y_water1m_n${N}_$tag.err
}
run b64 APX_DX_TRACE=1
run nobulk_c148 APX_DX_TRACE=1 APX_DX_BULK=0 APX_DX_CTAS=148
run b128 APX_DX_BULK_CTAS=128
for f in gpurun_out/r02y_water1m_n${N}_*.json; do
  python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "n_gpus", d["n_gpus"], "ms/step", round(d["ms_per_step"],3), "induce", round(d["ms_per_induce"],3), "iters", d["pcg_iterations"])
except Exception as e:
    print(sys.argv[1], "ERR", e)
PY
done
