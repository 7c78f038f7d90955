#!/bin/bash
# round 2, ninth GPU call: generic iteration graphs + deferred convergence check, list test beside the valence kernel, permanent
# field rows beside the PME round trip (and writing the pair tensors), 256-bit gathers
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | grep -v "^$" | tail -15 > gpurun_out/r02i_tests.log; tail -6 gpurun_out/r02i_tests.log
APX_TRACE_GRAPHS=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-strong > gpurun_out/r02i_dhfr2.json 2> gpurun_out/r02i_dhfr2.err
APX_LOOP=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-strong > gpurun_out/r02i_dhfr2_loop.json 2> gpurun_out/r02i_dhfr2_loop.err
timeout 300 python bench.py --workload water1m --mode energy --steps 5 --warmup 3 --no-cpu > gpurun_out/r02i_water1m.json 2> gpurun_out/r02i_water1m.err
timeout 300 python tools/trace_md.py --out gpurun_out/r02i_trace_md.txt > gpurun_out/r02i_trace_md.log 2>&1
for f in gpurun_out/r02i_dhfr2*.json gpurun_out/r02i_water1m*.json; do
  python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "ms/step", round(d["ms_per_step"],4), "induce", round(d["ms_per_induce"],4), "uf ms/launch", round(d["roofline"]["ms_per_launch"],4), "value", round(d["value"],2), "e2e", round(d["e2e"]["value"],2), "iters", d["pcg_iterations"], "batch", d.get("md",{}).get("batch",{}).get("value"), "steps", d.get("md",{}).get("ms_steps"))
except Exception as e:
    print(sys.argv[1], "ERR", e)
PY
done
grep -h "apx\]" gpurun_out/r02i_dhfr2.err | head -30
tail -n 3 gpurun_out/r02i_*.err | tail -30
head -30 gpurun_out/r02i_trace_md.log
