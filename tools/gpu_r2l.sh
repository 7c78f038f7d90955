#!/bin/bash
# round 2, twelfth GPU call: conditional (IF-node) energy epilogue behind the deferred solver batch + resume path
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_zgpu_6_tlist.py tests/test_gpu_parity.py tests/test_zgpu_2_md.py tests/test_predictor.py tests/test_zgpu_4_replicas.py -m gpu -q 2>&1 | grep -v "^$" | tail -12
APX_TRACE_GRAPHS=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-strong > gpurun_out/r02l_dhfr2.json 2> gpurun_out/r02l_dhfr2.err
timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu --no-strong > gpurun_out/r02l_dhfr2_40.json 2> gpurun_out/r02l_dhfr2_40.err
for f in gpurun_out/r02l_dhfr2*.json; do
  python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "ms/step", round(d["ms_per_step"],4), "induce", round(d["ms_per_induce"],4), "value", round(d["value"],2), "e2e", round(d["e2e"]["value"],2), "iters", d["pcg_iterations"], "batch", d.get("md",{}).get("batch",{}).get("value"), "misses", d["md"].get("solver_batch_misses"), "steps", d.get("md",{}).get("ms_steps"))
except Exception as e:
    print(sys.argv[1], "ERR", e)
PY
done
grep -h "apx\]" gpurun_out/r02l_dhfr2.err | head -30
tail -n 3 gpurun_out/r02l_*.err
