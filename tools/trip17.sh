#!/bin/bash
TAG=${1:-t17}
OUT=gpurun_out/r02/$TAG
mkdir -p $OUT
timeout 600 ncu --set full --import-source on --clock-control none -k regex:bowdb_match -c 1 -f -o $OUT/bowdb python tools/ncu_bowdb.py 2000 1 > $OUT/ncu.log 2>&1; tail -3 $OUT/ncu.log
