#!/bin/bash
# round 2: ncu launch list (gpu__time_duration) of the default bench command's MD steps
mkdir -p gpurun_out
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 2600 --csv --log-file gpurun_out/r02zz_md_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu --no-strong --no-ref-cuda > gpurun_out/r02zz_ncu.log 2>&1
echo "ncu rc=$?"; wc -l gpurun_out/r02zz_md_launches.csv; tail -2 gpurun_out/r02zz_ncu.log | cut -c1-300
python tools/summarize_launches.py gpurun_out/r02zz_md_launches.csv k_md_zero1 > gpurun_out/r02zz_md_launches.txt 2>&1; head -40 gpurun_out/r02zz_md_launches.txt
