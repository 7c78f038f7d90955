#!/bin/bash
# round 2, visit p (1 GPU): direct transport between two processes sharing the GPU, full GPU suite with the 1e-5 force tolerance
mkdir -p gpurun_out
nvidia-smi -L | head -2
timeout 500 python tools/direct_check.py --world 2 --blob water30 --timeout 400 > gpurun_out/r02p_direct_w2.log 2>&1
echo "direct w2 rc=$?"; grep -E "RESULT|Error|error" gpurun_out/r02p_direct_w2.log | head; tail -3 gpurun_out/r02p_direct_w2.log
timeout 500 python tools/direct_check.py --world 4 --blob water30 --timeout 400 > gpurun_out/r02p_direct_w4.log 2>&1
echo "direct w4 rc=$?"; grep -E "RESULT|Error|error" gpurun_out/r02p_direct_w4.log | head; tail -3 gpurun_out/r02p_direct_w4.log
timeout 1200 python -m pytest tests -q -m gpu --deselect tests/test_gpu_dist.py::test_direct_transport_between_processes > gpurun_out/r02p_tests.log 2>&1
echo "tests rc=$?"; tail -30 gpurun_out/r02p_tests.log
