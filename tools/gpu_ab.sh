#!/bin/bash
# A/B of an environment switch on one GPU: tests, then dhfr2 and water1m benches with and without it.
# usage: bash tools/gpu_ab.sh <tag> <ENVVAR> [pytest-args]
tag=${1:-ab}; var=${2:-APX_NO_RECORDS}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q ${3:-} > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/${tag}_pytest.log
for v in 0 1; do
  for wl in dhfr2 water1m; do
    steps=20; [ $wl = water1m ] && steps=5
    env $var=$v timeout 300 python bench.py --workload $wl --steps $steps --warmup 5 --no-cpu > gpurun_out/${tag}_${wl}_$v.json 2> gpurun_out/${tag}_${wl}_$v.err
    echo "$var=$v $wl rc=$? $(python -c "import json; d=json.loads(open('gpurun_out/${tag}_${wl}_$v.json').read().strip().splitlines()[-1]); print('ms/step', round(d['ms_per_step'],4), 'induce', round(d['ms_per_induce'],4), 'ufield launch', round(d['roofline']['ms_per_launch'],5))")"
  done
done
