#!/bin/bash
# One GPU-box visit: parity tests, bench line, ncu launch list of a short bench run.
# usage (under gpurun): bash tools/gpu_round.sh <tag> [pytest-args]
tag=${1:-run}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q ${2:-} > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest rc=$?"
tail -12 gpurun_out/${tag}_pytest.log
timeout 500 python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
echo "bench rc=$?"
cat gpurun_out/${tag}_bench.json
tail -5 gpurun_out/${tag}_bench.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/${tag}_ncu.log 2>&1
echo "ncu rc=$?"
