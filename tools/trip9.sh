#!/bin/bash
TAG=${1:-t9}
OUT=gpurun_out/r02/$TAG
mkdir -p $OUT
timeout 900 python bench.py --config 2 --steps 20 --warmup 3 > $OUT/bench_c2.json 2> $OUT/bench_c2.err; echo "config 2 rc=$?"; cat $OUT/bench_c2.json; tail -5 $OUT/bench_c2.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fast_kernel -s 6 -c 1 -o $OUT/fast_v4 python bench.py --config 1 --steps 1 --warmup 3 --no-cpu-baseline --no-parity > $OUT/fast_ncu.log 2>&1
ls -la $OUT
