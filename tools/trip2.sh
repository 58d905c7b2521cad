#!/bin/bash
TAG=${1:-t2}
OUT=gpurun_out/r02/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_kfdb.py -m gpu -x -q > $OUT/pytest_kfdb.log 2>&1; echo "rc=$?" >> $OUT/pytest_kfdb.log
tail -15 $OUT/pytest_kfdb.log
timeout 600 python tools/bench_configs.py > $OUT/configs.json 2> $OUT/configs.err; cat $OUT/configs.json; tail -3 $OUT/configs.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/match_launches.csv python tools/bench_configs.py --kfs 2000 --reps 3 > $OUT/match_ncu.log 2>&1
grep -E "bowdb|kfdb" $OUT/match_launches.csv | awk -F'","' '{print $5, $(NF-1), $NF}' | sort | uniq -c | sort -rn | head -20
