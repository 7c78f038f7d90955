#!/bin/bash
# First GPU call of a round for the reference-CUDA comparator and the drop-in (about 3 GPU-minutes on one B200):
#   gpurun --timeout 600 -- 'bash tools/gpu_refcuda.sh'
# 1. the three comparator / drop-in tests (all three passed their first runs in round 1, profiles/r01_refcuda*.json, r01_dropin*.json);
# 2. ours vs the reference's CUDA build in one job, same conditions, on dhfr2, the 96k water box and the 424k protein box;
# 3. the reference's launch list for one induce() + one energy step on dhfr2 (ncu, times cold-cache: shares only).
mkdir -p gpurun_out
python -m pytest tests/test_zgpu_9_refcuda.py tests/test_polpair_golden.py -m gpu -q -rxXs > gpurun_out/refcuda_tests.log 2>&1      # polpair: XPASS = remove its xfail marker
tail -5 gpurun_out/refcuda_tests.log
for w in dhfr2 water96k dhfr424k; do
   timeout 300 python tools/same_job_compare.py 30 $w > gpurun_out/samejob_$w.json 2> gpurun_out/samejob_$w.err
   echo "$w: $(head -c 400 gpurun_out/samejob_$w.json)"
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/refcuda_launches.csv \
   python -m oracle.ref_cuda_bridge tests/golden/dhfr2.npz --reps 1 --warmup 0 > gpurun_out/refcuda_ncu.log 2>&1
python tools/summarize_launches.py gpurun_out/refcuda_launches.csv > gpurun_out/refcuda_launches.txt 2>/dev/null
head -30 gpurun_out/refcuda_launches.txt
