#!/bin/bash
# round 2, eighth GPU call: device-side PCG loop (conditional WHILE node), 256-bit gathers, operator CTA cap
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_zgpu_6_tlist.py tests/test_gpu_parity.py tests/test_predictor.py tests/test_zgpu_2_md.py tests/test_zgpu_3_rebuild.py -m gpu -q -x 2>&1 | grep -v "^$" | tail -15
APX_NO_LOOP=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | grep -v "^$" | tail -3
for ct in 6 3 12; do
  APX_TRACE_GRAPHS=1 APX_TL_CTAS=$ct timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-strong > gpurun_out/r02h_dhfr2_ct$ct.json 2> gpurun_out/r02h_dhfr2_ct$ct.err
done
APX_NO_LOOP=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-strong > gpurun_out/r02h_dhfr2_noloop.json 2> gpurun_out/r02h_dhfr2_noloop.err
timeout 300 python bench.py --workload water1m --mode energy --steps 5 --warmup 3 --no-cpu > gpurun_out/r02h_water1m.json 2> gpurun_out/r02h_water1m.err
timeout 300 python tools/trace_md.py --out gpurun_out/r02h_trace_md.txt > gpurun_out/r02h_trace_md.log 2>&1
for f in gpurun_out/r02h_dhfr2_*.json gpurun_out/r02h_water1m*.json; do
  python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "ms/step", round(d["ms_per_step"],4), "induce", round(d["ms_per_induce"],4), "uf ms/launch", round(d["roofline"]["ms_per_launch"],4), "value", round(d["value"],2), "iters", d["pcg_iterations"], "steps", d.get("md",{}).get("ms_steps"))
except Exception as e:
    print(sys.argv[1], "ERR", e)
PY
done
grep -h "apx\]" gpurun_out/r02h_dhfr2_ct6.err | head -30
tail -n 3 gpurun_out/r02h_*.err | tail -30
head -30 gpurun_out/r02h_trace_md.log
