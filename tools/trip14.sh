#!/bin/bash
TAG=${1:-t14}
OUT=gpurun_out/r02/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_kfdb.py tests/test_gpu_match.py tests/test_gpu_adapters.py tests/test_gpu_voc_real.py -m gpu -x -q > $OUT/pytest.log 2>&1; echo "rc=$?" >> $OUT/pytest.log
tail -8 $OUT/pytest.log
timeout 900 python bench.py --config 4 --steps 20 --warmup 3 --no-cpu-baseline > $OUT/bench_c4.json 2> $OUT/bench_c4.err; echo "config 4 rc=$?"; cat $OUT/bench_c4.json; tail -3 $OUT/bench_c4.err
