#!/bin/bash
# round 2, tenth GPU call: whole suite; WHILE-node loop as default; stored-tensor preconditioner; one-sweep row compaction;
# second-generation dipole spread / gather (A/B)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | grep -v "^$" | tail -25 > gpurun_out/r02j_tests.log; tail -8 gpurun_out/r02j_tests.log
APX_TRACE_GRAPHS=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-strong > gpurun_out/r02j_dhfr2.json 2> gpurun_out/r02j_dhfr2.err
APX_PME_GEN2=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-strong > gpurun_out/r02j_dhfr2_gen1.json 2> gpurun_out/r02j_dhfr2_gen1.err
APX_LOOP=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-strong > gpurun_out/r02j_dhfr2_generic.json 2> gpurun_out/r02j_dhfr2_generic.err
timeout 300 python bench.py --workload water1m --mode energy --steps 5 --warmup 3 --no-cpu > gpurun_out/r02j_water1m.json 2> gpurun_out/r02j_water1m.err
APX_PME_GEN2=0 timeout 300 python bench.py --workload water1m --mode energy --steps 5 --warmup 3 --no-cpu > gpurun_out/r02j_water1m_gen1.json 2> gpurun_out/r02j_water1m_gen1.err
APX_LOOP=0 timeout 300 python tools/trace_md.py --out gpurun_out/r02j_trace_md.txt > gpurun_out/r02j_trace_md.log 2>&1
APX_LOOP=0 timeout 300 python tools/trace_step.py --workload water1m --steps 2 --out gpurun_out/r02j_trace_water1m.txt > gpurun_out/r02j_trace_water1m.log 2>&1
for f in gpurun_out/r02j_dhfr2*.json gpurun_out/r02j_water1m*.json; do
  python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "ms/step", round(d["ms_per_step"],4), "induce", round(d["ms_per_induce"],4), "uf ms/launch", round(d["roofline"]["ms_per_launch"],4), "value", round(d["value"],2), "e2e", round(d["e2e"]["value"],2), "iters", d["pcg_iterations"], "batch", d.get("md",{}).get("batch",{}).get("value"), "steps", d.get("md",{}).get("ms_steps"))
except Exception as e:
    print(sys.argv[1], "ERR", e)
PY
done
grep -h "apx\]" gpurun_out/r02j_dhfr2.err | head -30
tail -n 3 gpurun_out/r02j_*.err | tail -30
head -32 gpurun_out/r02j_trace_md.log
head -40 gpurun_out/r02j_trace_water1m.txt
