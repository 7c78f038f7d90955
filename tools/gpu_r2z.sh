#!/bin/bash
# round 2, visit z (gpurun --gpus 8): operator forked after the spread in decomposed runs (APX_DIST_FORK_LATE), with and without a cap
N=8
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
run() { # tag, env...
  tag=$1; shift
  env "$@" timeout 300 $TR --master-port 29613 bench.py --gpus $N --workload water1m --steps 5 --warmup 3 --no-cpu > gpurun_out/r02z_water1m_n${N}_$tag.json 2> gpurun_out/r02z_water1m_n${N}_$tag.err
  echo "water1m N=$N $tag rc=$?"; grep "apx dx trace\] rank 0" gpurun_out/r02z_water1m_n${N}_$tag.err
}
run late APX_DX_TRACE=1
run late_cap6 APX_DX_TRACE=1 APX_TL_CTAS=6
for f in gpurun_out/r02z_water1m_n${N}_*.json; do
  python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "n_gpus", d["n_gpus"], "ms/step", round(d["ms_per_step"],3), "induce", round(d["ms_per_induce"],3), "iters", d["pcg_iterations"])
except Exception as e:
    print(sys.argv[1], "ERR", e)
PY
done
