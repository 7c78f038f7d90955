#!/bin/bash
# round 2, third GPU call: precision diagnostics after the cell-vector fix, the whole GPU suite (new: device-pointer entry points,
# replica parity of the large configurations, drop-in through emplar / energyReduce), the default bench line with its
# strong-scaling leg at N = 1
mkdir -p gpurun_out
timeout 400 python tools/diag_precision.py water30 dhfr2 > gpurun_out/r02c_diag.log 2>&1
timeout 1200 python -m pytest tests -m gpu -q -s 2>&1 | grep -v "^$" | tail -60 > gpurun_out/r02c_tests.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02c_bench.json 2> gpurun_out/r02c_bench.err
cat gpurun_out/r02c_diag.log
grep -E "passed|failed|FAILED|Error|water|dhfr424k|reference front" gpurun_out/r02c_tests.log | cut -c1-900
tail -c 3000 gpurun_out/r02c_bench.json; tail -5 gpurun_out/r02c_bench.err
