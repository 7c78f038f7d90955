#!/bin/bash
# round 2, visit v (1 GPU): full GPU suite (fixed-point spreading test, process-per-rank direct transport test), bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r02v_tests.log 2>&1
echo "tests rc=$?"; tail -15 gpurun_out/r02v_tests.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02v_bench.json 2> gpurun_out/r02v_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/r02v_bench.err; python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02v_bench.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], "roofline", d["roofline"]["frac"], "strong", d.get("strong_scaling",{}).get("ms_per_step"))
PY
APX_PME_FIXED=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --no-strong > gpurun_out/r02v_bench_fixed.json 2> gpurun_out/r02v_bench_fixed.err
echo "bench fixed rc=$?"; python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02v_bench_fixed.json").read().strip().splitlines()[-1])
print("fixed-point spreading: value", d["value"], "ms/step", d["ms_per_step"], "induce", d["ms_per_induce"])
PY
