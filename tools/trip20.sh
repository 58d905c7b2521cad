#!/bin/bash
TAG=${1:-t20}
OUT=gpurun_out/r02/$TAG
mkdir -p $OUT
for c in 4 2 1; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file $OUT/launches_c$c.csv python bench.py --config $c --steps 2 --warmup 1 --no-cpu-baseline --no-parity > $OUT/ncu_c$c.log 2>&1
  echo "config $c rc=$? lines=$(wc -l < $OUT/launches_c$c.csv)"
done
