#!/bin/bash
# round 2, GPU call: whole suite + the default bench line exactly as the driver runs it (+ the reference arm) after the epilogue fork
mkdir -p gpurun_out
nproc
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | grep -v "^$" | tail -12 > gpurun_out/r02n_tests.log; tail -5 gpurun_out/r02n_tests.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02n_bench.json 2> gpurun_out/r02n_bench.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02n_bench_reference.json 2> gpurun_out/r02n_bench_reference.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02n_bench.json").read().strip().splitlines()[-1])
print("value", round(d["value"],2), "ms/step", round(d["ms_per_step"],4), "induce", round(d["ms_per_induce"],4), "e2e", round(d["e2e"]["value"],2), "batch", d["md"]["batch"]["value"], "misses", d["md"].get("solver_batch_misses"))
print("steps", d["md"]["ms_steps"])
print("roofline", json.dumps(d["roofline"])[:900])
print("cpu_baseline", json.dumps(d.get("cpu_baseline"))[:600])
print("ref_cuda", json.dumps(d.get("ref_cuda"))[:1500])
print("strong", json.dumps(d.get("strong_scaling"))[:800])
r=json.loads(open("gpurun_out/r02n_bench_reference.json").read().strip().splitlines()[-1])
print("reference arm", json.dumps(r)[:1200])
PY
tail -n 3 gpurun_out/r02n_*.err
