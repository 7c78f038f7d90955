#!/bin/bash
# round 2, second GPU call: where the mixed build's force error comes from + the new tests (device-pointer entry points,
# replica parity of the large configurations, the drop-in through the reference's emplar / energyReduce)
mkdir -p gpurun_out
timeout 600 python tools/diag_precision.py water30 dhfr2 > gpurun_out/r02b_diag.log 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_zgpu_4_replicas.py tests/test_zgpu_9_refcuda.py -m gpu -x -q -s 2>&1 | tail -40 > gpurun_out/r02b_tests.log
cat gpurun_out/r02b_diag.log
tail -30 gpurun_out/r02b_tests.log
