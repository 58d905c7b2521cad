#!/bin/bash
TAG=${1:-t6}
OUT=gpurun_out/r02/$TAG
mkdir -p $OUT
for c in 2 4 1; do
  timeout 900 python bench.py --config $c --steps 20 --warmup 3 > $OUT/bench_c$c.json 2> $OUT/bench_c$c.err; echo "config $c rc=$?"; cat $OUT/bench_c$c.json; tail -5 $OUT/bench_c$c.err
done
timeout 300 python bench.py --impl reference --config 2 --steps 3 --warmup 1 > $OUT/ref_c2.json 2> $OUT/ref_c2.err; cat $OUT/ref_c2.json | cut -c1-400
timeout 300 python bench.py --impl reference --config 4 --steps 3 --warmup 1 > $OUT/ref_c4.json 2> $OUT/ref_c4.err; cat $OUT/ref_c4.json | cut -c1-400
