#!/bin/bash
# round 2, visit t (gpurun --gpus N): exchange-kernel timestamps (APX_DX_TRACE) and CTA sweep of the direct transport
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
run() { # tag, env...
  tag=$1; shift
  env "$@" timeout 400 $TR --master-port 29613 bench.py --gpus $N --workload water1m --steps 5 --warmup 3 --no-cpu > gpurun_out/r02t_water1m_n${N}_$tag.json 2> gpurun_out/r02t_water1m_n${N}_$tag.err
  echo "water1m N=$N $tag rc=$?"; grep "apx dx trace\] rank 0" gpurun_out/r02t_water1m_n${N}_$tag.err
}
run c296 APX_DX_TRACE=1
run c64 APX_DX_TRACE=1 APX_DX_CTAS=64
run c32 APX_DX_TRACE=1 APX_DX_CTAS=32
for f in gpurun_out/r02t_water1m_n${N}_*.json; do
  python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "n_gpus", d["n_gpus"], "ms/step", round(d["ms_per_step"],3), "induce", round(d["ms_per_induce"],3), "iters", d["pcg_iterations"])
except Exception as e:
    print(sys.argv[1], "ERR", e)
PY
done
