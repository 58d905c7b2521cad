#!/bin/bash
# ncu --set full of one steady-state pass of config 1 and config 2 (single handle), and of the database query kernels
TAG=${1:-t35}
OUT=gpurun_out/r02/$TAG
mkdir -p $OUT
timeout 900 ncu --set full --clock-control none --launch-skip 170 -c 17 -f -o $OUT/step_c1 python bench.py --config 1 --steps 2 --warmup 1 --handles 1 --no-parity --no-cpu-baseline > $OUT/ncu_c1.log 2>&1; echo "c1 rc=$?"
timeout 900 ncu --set full --clock-control none --launch-skip 220 -c 18 -f -o $OUT/step_c2 python bench.py --config 2 --steps 2 --warmup 1 --handles 1 --no-parity --no-cpu-baseline > $OUT/ncu_c2.log 2>&1; echo "c2 rc=$?"
timeout 600 ncu --set full --clock-control none -k regex:"bowdb|kfdb_score|bow_transform" --launch-skip 2020 -c 4 -f -o $OUT/query_c4 python tools/ncu_bowdb.py 2000 1 > $OUT/ncu_c4.log 2>&1; echo "c4 rc=$?"
ls -la $OUT
