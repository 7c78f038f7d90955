#!/bin/bash
# round 2, final visit (1 GPU): smoke, full GPU suite, the driver's bench line and its reference arm
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02zz_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r02zz_smoke.log
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r02zz_tests.log 2>&1
echo "tests rc=$?"; tail -6 gpurun_out/r02zz_tests.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02zz_bench.json 2> gpurun_out/r02zz_bench.err
echo "bench rc=$?"; tail -2 gpurun_out/r02zz_bench.err; python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02zz_bench.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], "roofline", d["roofline"]["frac"], "strong", d.get("strong_scaling",{}).get("ms_per_step"), "launches", d["gpu_launches"])
PY
timeout 600 python bench.py --impl reference --steps 2 --warmup 0 > gpurun_out/r02zz_bench_reference.json 2> gpurun_out/r02zz_bench_reference.err
echo "reference arm rc=$?"; tail -c 600 gpurun_out/r02zz_bench_reference.json
