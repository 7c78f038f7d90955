#!/bin/bash
# round 2, seventh GPU call: conditional-graph probe, A/B after the launch-order / tensor-flag / L2-hint fixes, 1 M-atom timeline,
# which buffer moves at the first MD rebuild
mkdir -p gpurun_out
./tools/probe/cond_probe > gpurun_out/r02g_cond_probe.txt 2>&1; cat gpurun_out/r02g_cond_probe.txt
timeout 600 python -m pytest tests/test_zgpu_6_tlist.py tests/test_gpu_parity.py -m gpu -q -x 2>&1 | grep -v "^$" | tail -5
for m in 1 0; do
  APX_TRACE_GRAPHS=1 APX_TL_MODE=$m timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-strong > gpurun_out/r02g_dhfr2_tl$m.json 2> gpurun_out/r02g_dhfr2_tl$m.err
done
APX_TLIST=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-strong > gpurun_out/r02g_dhfr2_rows.json 2> gpurun_out/r02g_dhfr2_rows.err
APX_TL_MODE=1 timeout 300 python bench.py --workload water1m --mode energy --steps 5 --warmup 3 --no-cpu > gpurun_out/r02g_water1m_tl1.json 2> gpurun_out/r02g_water1m_tl1.err
timeout 300 python tools/trace_step.py --workload water1m --steps 2 --out gpurun_out/r02g_trace_water1m.txt > gpurun_out/r02g_trace_water1m.log 2>&1
timeout 300 python tools/trace_md.py --out gpurun_out/r02g_trace_md.txt > gpurun_out/r02g_trace_md.log 2>&1
for f in gpurun_out/r02g_dhfr2_*.json gpurun_out/r02g_water1m_*.json; do
  python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "ms/step", round(d["ms_per_step"],4), "induce", round(d["ms_per_induce"],4), "uf ms/launch", round(d["roofline"]["ms_per_launch"],4), "value", round(d["value"],2), "steps", d.get("md",{}).get("ms_steps"))
except Exception as e:
    print(sys.argv[1], "ERR", e)
PY
done
grep -h "apx\]" gpurun_out/r02g_dhfr2_tl1.err | head -30
head -45 gpurun_out/r02g_trace_water1m.txt
