#!/bin/bash
TAG=${1:-t15}
OUT=gpurun_out/r02/$TAG
mkdir -p $OUT
timeout 300 python tools/ncu_bowdb.py 2000 8 > $OUT/bowdb_times.json 2> $OUT/bowdb_times.err; cat $OUT/bowdb_times.json; tail -3 $OUT/bowdb_times.err
timeout 600 ncu --set full --import-source on --clock-control none -k regex:bowdb -c 2 -f -o $OUT/bowdb python tools/ncu_bowdb.py 2000 1 > $OUT/ncu.log 2>&1; tail -3 $OUT/ncu.log
ls -la $OUT
