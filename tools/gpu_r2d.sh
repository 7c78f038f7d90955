#!/bin/bash
# round 2, fourth GPU call: the staged operator (tests, A/B against the row operator on dhfr2 and the 1 M-atom box, ncu),
# precision diagnostics with the rebuilt variants, the tests that failed in call 3
mkdir -p gpurun_out
timeout 300 python tools/diag_precision.py water30 dhfr2 > gpurun_out/r02d_diag.log 2>&1
timeout 900 python -m pytest tests/test_zgpu_5_staged.py tests/test_gpu_parity.py tests/test_zgpu_9_refcuda.py tests/test_zgpu_4_replicas.py -m gpu -q -s 2>&1 | grep -v "^$" | tail -60 > gpurun_out/r02d_tests.log
for st in 1 0; do
  APX_STAGED=$st timeout 300 python bench.py --workload water1m --mode energy --steps 5 --warmup 3 --no-cpu > gpurun_out/r02d_water1m_staged$st.json 2> gpurun_out/r02d_water1m_staged$st.err
  APX_STAGED=$st timeout 300 python bench.py --steps 20 --warmup 8 --no-cpu --no-strong > gpurun_out/r02d_dhfr2_staged$st.json 2> gpurun_out/r02d_dhfr2_staged$st.err
done
for cap in 40 48; do
  APX_STAGED_CAP=$cap timeout 300 python bench.py --workload water1m --mode energy --steps 5 --warmup 3 --no-cpu > gpurun_out/r02d_water1m_cap$cap.json 2> gpurun_out/r02d_water1m_cap$cap.err
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_ufield_staged|k_rows_compact_grp|k_mplar_listed|k_mplar_rows|k_precond_rows" -s 6 -c 6 -f -o gpurun_out/r02d_water1m \
   python bench.py --workload water1m --mode energy --steps 1 --warmup 3 --no-cpu > gpurun_out/r02d_ncu.log 2>&1
cat gpurun_out/r02d_diag.log
grep -E "passed|failed|FAILED|Error|^water|^dhfr424k|reference front" gpurun_out/r02d_tests.log | cut -c1-1200
for f in gpurun_out/r02d_water1m_staged1.json gpurun_out/r02d_water1m_staged0.json gpurun_out/r02d_water1m_cap40.json gpurun_out/r02d_water1m_cap48.json gpurun_out/r02d_dhfr2_staged1.json gpurun_out/r02d_dhfr2_staged0.json; do
  python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print(sys.argv[1], "ms/step", round(d["ms_per_step"],4), "induce", round(d["ms_per_induce"],4), "uf ms/launch", d["roofline"]["ms_per_launch"], "frac", round(d["roofline"]["frac"],4), "value", round(d["value"],2), "md" , json.dumps(d.get("md",{}))[:300])
except Exception as e:
    print(sys.argv[1], "ERR", e)
PY
done
tail -3 gpurun_out/r02d_*.err | tail -30
