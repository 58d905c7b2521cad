#!/bin/bash
TAG=${1:-t3}
OUT=gpurun_out/r02/$TAG
mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1; echo "rc=$?" >> $OUT/pytest.log
tail -25 $OUT/pytest.log
timeout 600 python tools/bench_configs.py > $OUT/configs.json 2> $OUT/configs.err; cat $OUT/configs.json; tail -3 $OUT/configs.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/match_launches.csv python tools/bench_configs.py --kfs 200 --reps 3 > $OUT/match_ncu.log 2>&1
grep -E "proj_|grid_sort|project_points" $OUT/match_launches.csv | awk -F'","' '{print $5, $(NF-1), $NF}' | sort | uniq -c | sort -rn | head -20
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bowdb_match_kernel -c 2 -o $OUT/bowdb_v1 python tools/bench_configs.py --kfs 2000 --reps 2 > $OUT/bow_ncu.log 2>&1
