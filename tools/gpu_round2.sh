#!/bin/bash
# Full single-GPU visit: whole GPU suite, default bench, bench with the vdW term, launch lists, ncu --set full of the top kernels.
tag=${1:-run}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest rc=$?"; tail -6 gpurun_out/${tag}_pytest.log
timeout 500 python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
echo "bench rc=$?"; cat gpurun_out/${tag}_bench.json; tail -3 gpurun_out/${tag}_bench.err
timeout 300 python bench.py --steps 20 --warmup 5 --vdw --no-cpu > gpurun_out/${tag}_bench_vdw.json 2> gpurun_out/${tag}_bench_vdw.err
echo "bench vdw rc=$?"; cat gpurun_out/${tag}_bench_vdw.json; tail -3 gpurun_out/${tag}_bench_vdw.err
timeout 300 python bench.py --workload water1m --steps 5 --warmup 3 --vdw > gpurun_out/${tag}_water1m_vdw.json 2> gpurun_out/${tag}_water1m_vdw.err
echo "water1m vdw rc=$?"; cat gpurun_out/${tag}_water1m_vdw.json; tail -3 gpurun_out/${tag}_water1m_vdw.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --vdw > gpurun_out/${tag}_ncu.log 2>&1
echo "ncu rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_ufield_rows|k_ehal_rows|k_spread_dp|k_gather_dp|k_mplar_rows|k_precond_rows" -s 30 -c 12 -f -o gpurun_out/${tag}_full \
   python bench.py --workload water1m --steps 1 --warmup 3 --vdw > gpurun_out/${tag}_ncu_full.log 2>&1
echo "ncu full rc=$?"; tail -2 gpurun_out/${tag}_ncu_full.log
