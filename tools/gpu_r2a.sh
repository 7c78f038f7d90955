#!/bin/bash
# round 2, first GPU call: the whole GPU suite after the coordinate / listed-pair changes + the per-operator differences
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r02a_gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r02a_tests.log
timeout 600 python tests/gpu_check.py water dhfr > gpurun_out/r02a_check.log 2>&1
tail -5 gpurun_out/r02a_tests.log
grep -E "==|grad|induce|energy|virial" gpurun_out/r02a_check.log
