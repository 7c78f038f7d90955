#!/bin/bash
# round 2, visit x (1 GPU): vdW scalars copied after the merged epilogue, report copy in front of the step's only synchronisation
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_zgpu_2_md.py tests/test_zgpu_3_rebuild.py tests/test_drivers.py tests/test_gpu_vdw.py -q -m gpu -x > gpurun_out/r02x_tests.log 2>&1
echo "tests rc=$?"; tail -5 gpurun_out/r02x_tests.log
timeout 600 python bench.py --steps 40 --warmup 8 --no-cpu --no-strong > gpurun_out/r02x_dhfr2.json 2> gpurun_out/r02x_dhfr2.err
python - gpurun_out/r02x_dhfr2.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], "value", round(d["value"],2), "ms/step", round(d["ms_per_step"],4), "median", round(d["md"]["ms_per_step_median"],4), "e2e", round(d["e2e"]["value"],2), "batch", round(d["md"]["batch"]["value"],2), "misses", d["md"].get("solver_batch_misses"), "wall", d.get("wall_s_timed_region"))
PY
timeout 300 python tools/trace_md.py --steps 24 --out gpurun_out/r02x_trace_md.txt > gpurun_out/r02x_trace_md.log 2>&1; head -14 gpurun_out/r02x_trace_md.txt
