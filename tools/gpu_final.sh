#!/bin/bash
# End-of-round visit on one GPU: whole GPU suite, smoke, default bench (with cpu_baseline), bench with vdW, launch list, ncu --set full (dhfr2).
tag=${1:-final}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/${tag}_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"; grep smoke: gpurun_out/${tag}_smoke.log
timeout 500 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
echo "bench rc=$?"; cat gpurun_out/${tag}_bench.json; tail -3 gpurun_out/${tag}_bench.err
timeout 300 python bench.py --vdw --no-cpu > gpurun_out/${tag}_bench_vdw.json 2> gpurun_out/${tag}_bench_vdw.err
echo "bench vdw rc=$?"; cat gpurun_out/${tag}_bench_vdw.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --vdw > gpurun_out/${tag}_ncu.log 2>&1
echo "ncu rc=$?"
APX_NO_GRAPH=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_ufield_rows|k_ehal_rows|k_spread_dp|k_gather_dp|k_mplar_rows|k_precond_rows|k_fft64" -s 40 -c 14 -f -o gpurun_out/${tag}_full \
   python bench.py --steps 1 --warmup 3 --vdw --no-cpu > gpurun_out/${tag}_ncu_full.log 2>&1
echo "ncu full rc=$?"; tail -2 gpurun_out/${tag}_ncu_full.log
