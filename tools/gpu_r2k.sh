#!/bin/bash
# round 2, eleventh GPU call: whole suite with the final defaults, vdW fork point A/B, launch lists and ncu --set full of the new
# kernels at dhfr2 (MD step) and at 1 M atoms
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | grep -v "^$" | tail -25 > gpurun_out/r02k_tests.log; tail -6 gpurun_out/r02k_tests.log
for at in 1 0 2; do
  APX_VDW_AT=$at APX_TRACE_GRAPHS=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-strong > gpurun_out/r02k_dhfr2_vdw$at.json 2> gpurun_out/r02k_dhfr2_vdw$at.err
done
timeout 300 python bench.py --workload water1m --mode energy --steps 5 --warmup 3 --no-cpu > gpurun_out/r02k_water1m.json 2> gpurun_out/r02k_water1m.err
timeout 300 python tools/trace_md.py --out gpurun_out/r02k_trace_md.txt > gpurun_out/r02k_trace_md.log 2>&1
for f in gpurun_out/r02k_dhfr2*.json gpurun_out/r02k_water1m*.json; do
  python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "ms/step", round(d["ms_per_step"],4), "induce", round(d["ms_per_induce"],4), "uf ms/launch", round(d["roofline"]["ms_per_launch"],4), "value", round(d["value"],2), "e2e", round(d["e2e"]["value"],2), "iters", d["pcg_iterations"], "batch", d.get("md",{}).get("batch",{}).get("value"), "steps", d.get("md",{}).get("ms_steps"))
except Exception as e:
    print(sys.argv[1], "ERR", e)
PY
done
grep -h "apx\]" gpurun_out/r02k_dhfr2_vdw1.err | head -30
head -32 gpurun_out/r02k_trace_md.log
# launch lists (ncu, one metric) and --set full captures
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02k_md_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-strong --no-ref-cuda > gpurun_out/r02k_md_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_ufield_tl|k_spread_dp2|k_gather_dp2|k_precond_rows|k_rows_compact|k_dfield_rows|k_mplar_rows|k_ehal_rows" -s 40 -c 16 -f -o gpurun_out/r02k_dhfr2 python bench.py --steps 1 --warmup 3 --no-cpu --no-strong --no-ref-cuda > gpurun_out/r02k_ncu_dhfr2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_rows_build" -c 4 -f -o gpurun_out/r02k_dhfr2_build python bench.py --steps 1 --warmup 3 --no-cpu --no-strong --no-ref-cuda > gpurun_out/r02k_ncu_dhfr2_build.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_ufield_tl|k_spread_dp2|k_gather_dp2|k_precond_rows|k_dfield_rows|k_mplar_rows" -s 12 -c 10 -f -o gpurun_out/r02k_water1m python bench.py --workload water1m --mode energy --steps 1 --warmup 3 --no-cpu > gpurun_out/r02k_ncu_water1m.log 2>&1
ls -la gpurun_out/r02k_*.ncu-rep
tail -n 2 gpurun_out/r02k_ncu_*.log | tail -20
