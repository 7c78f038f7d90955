#!/bin/bash
# Large-configuration visit (BASELINE.json configs[2..4]) on ONE GPU: bench lines + ncu launch list of the 1 M-atom box.
# usage (under gpurun): bash tools/gpu_big.sh <tag> [pytest-file]
tag=${1:-big}
mkdir -p gpurun_out
if [ -n "$2" ]; then
  timeout 900 python -m pytest $2 -m gpu -q -x > gpurun_out/${tag}_pytest.log 2>&1
  echo "pytest rc=$?"; tail -5 gpurun_out/${tag}_pytest.log
fi
for wl in water96k dhfr424k water1m; do
  timeout 400 python bench.py --workload $wl --steps 5 --warmup 3 > gpurun_out/${tag}_${wl}.json 2> gpurun_out/${tag}_${wl}.err
  echo "$wl rc=$?"; cat gpurun_out/${tag}_${wl}.json; tail -3 gpurun_out/${tag}_${wl}.err
done
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/${tag}_water1m_launches.csv \
    python bench.py --workload water1m --steps 1 --warmup 3 > gpurun_out/${tag}_ncu.log 2>&1
echo "ncu rc=$?"
