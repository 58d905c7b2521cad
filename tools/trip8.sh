#!/bin/bash
TAG=${1:-t8}
OUT=gpurun_out/r02/$TAG
mkdir -p $OUT
timeout 900 python bench.py --config 2 --steps 20 --warmup 3 > $OUT/bench_c2.json 2> $OUT/bench_c2.err; echo "config 2 rc=$?"; cat $OUT/bench_c2.json; tail -5 $OUT/bench_c2.err
timeout 900 python -m pytest tests/test_gpu_voc_real.py -m gpu -x -q -s > $OUT/pytest_voc.log 2>&1; tail -6 $OUT/pytest_voc.log
