#!/bin/bash
# round 2, GPU call: MD tail behind the deferred energy evaluation (one synchronisation per step), in-iteration operator timing,
# cheaper vdW reciprocal; MD + vdW + parity tests, bench, trace
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_zgpu_2_md.py tests/test_gpu_vdw.py tests/test_zgpu_6_tlist.py tests/test_gpu_parity.py tests/test_drivers.py tests/test_zgpu_3_rebuild.py -m gpu -q 2>&1 | grep -v "^$" | tail -12
APX_TRACE_GRAPHS=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-strong > gpurun_out/r02o_dhfr2.json 2> gpurun_out/r02o_dhfr2.err
timeout 300 python bench.py --steps 60 --warmup 5 --no-cpu --no-strong > gpurun_out/r02o_dhfr2_60.json 2> gpurun_out/r02o_dhfr2_60.err
timeout 300 python tools/trace_md.py --out gpurun_out/r02o_trace_md.txt > gpurun_out/r02o_trace_md.log 2>&1
for f in gpurun_out/r02o_dhfr2*.json; do
  python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "ms/step", round(d["ms_per_step"],4), "median", round(d["md"]["ms_per_step_median"],4), "induce", round(d["ms_per_induce"],4), "value", round(d["value"],2), "e2e", round(d["e2e"]["value"],2), "batch", d["md"]["batch"]["value"], "misses", d["md"].get("solver_batch_misses"), "uf", d["roofline"]["ms_per_launch"], "frac", d["roofline"]["frac"], "T", d["md"]["temperature_K"])
    print(d["md"]["ms_steps"])
except Exception as e:
    print(sys.argv[1], "ERR", e)
PY
done
grep -h "apx\]" gpurun_out/r02o_dhfr2.err | head -30
tail -n 3 gpurun_out/r02o_*.err
head -30 gpurun_out/r02o_trace_md.log
