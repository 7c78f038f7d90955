#!/bin/bash
# round 2, sixth GPU call: stored-tensor operator + Hilbert order + cooperative list build: tests, A/B, traces
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_zgpu_6_tlist.py tests/test_zgpu_3_rebuild.py tests/test_gpu_parity.py tests/test_gpu_vdw.py tests/test_gpu_dist.py -m gpu -q -x 2>&1 | grep -v "^$" | tail -30 > gpurun_out/r02f_tests.log
tail -5 gpurun_out/r02f_tests.log
timeout 300 python tools/trace_md.py --out gpurun_out/r02f_trace_md.txt > gpurun_out/r02f_trace_md.log 2>&1
head -30 gpurun_out/r02f_trace_md.log
for m in 0 1 2 3; do
  APX_TL_MODE=$m timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-strong > gpurun_out/r02f_dhfr2_tl$m.json 2> gpurun_out/r02f_dhfr2_tl$m.err
  APX_TL_MODE=$m timeout 300 python bench.py --workload water1m --mode energy --steps 5 --warmup 3 --no-cpu > gpurun_out/r02f_water1m_tl$m.json 2> gpurun_out/r02f_water1m_tl$m.err
done
APX_TLIST=0 APX_TRACE_GRAPHS=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-strong > gpurun_out/r02f_dhfr2_rows.json 2> gpurun_out/r02f_dhfr2_rows.err
APX_TLIST=0 timeout 300 python bench.py --workload water1m --mode energy --steps 5 --warmup 3 --no-cpu > gpurun_out/r02f_water1m_rows.json 2> gpurun_out/r02f_water1m_rows.err
for f in gpurun_out/r02f_dhfr2_*.json gpurun_out/r02f_water1m_*.json; do
  python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "ms/step", round(d["ms_per_step"],4), "induce", round(d["ms_per_induce"],4), "uf ms/launch", round(d["roofline"]["ms_per_launch"],4), "value", round(d["value"],2), "steps", d.get("md",{}).get("ms_steps"))
except Exception as e:
    print(sys.argv[1], "ERR", e)
PY
done
grep -h "apx\]" gpurun_out/r02f_dhfr2_rows.err | head -30
tail -3 gpurun_out/r02f_*.err | tail -20
