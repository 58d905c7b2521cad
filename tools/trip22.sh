#!/bin/bash
TAG=${1:-t22}
OUT=gpurun_out/r02/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_kfdb.py tests/test_gpu_adapters.py tests/test_gpu_voc_real.py -m gpu -x -q > $OUT/pytest.log 2>&1; echo "rc=$?" >> $OUT/pytest.log
tail -4 $OUT/pytest.log
timeout 600 python tools/bench_configs.py > $OUT/configs.json 2> $OUT/configs.err; tail -1 $OUT/configs.json; tail -3 $OUT/configs.err
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"bowdb|kfdb_score" -c 3 -f -o $OUT/kfdb python tools/ncu_bowdb.py 2000 1 > $OUT/ncu.log 2>&1; tail -2 $OUT/ncu.log
