#!/bin/bash
# round 2, fifth GPU call: whole GPU suite, MD step timeline (ordinary and rebuild steps), default bench line
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | grep -v "^$" | tail -30 > gpurun_out/r02e_tests.log
timeout 300 python tools/trace_md.py --out gpurun_out/r02e_trace_md.txt > gpurun_out/r02e_trace_md.log 2>&1
APX_STAGED=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --no-strong > gpurun_out/r02e_bench.json 2> gpurun_out/r02e_bench.err
tail -5 gpurun_out/r02e_tests.log
head -30 gpurun_out/r02e_trace_md.log
tail -3 gpurun_out/r02e_bench.err
