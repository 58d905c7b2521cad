#!/bin/bash
TAG=${1:-t7}
OUT=gpurun_out/r02/$TAG
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1; echo "rc=$?" >> $OUT/pytest.log
tail -15 $OUT/pytest.log
timeout 900 python bench.py --config 2 --steps 20 --warmup 3 > $OUT/bench_c2.json 2> $OUT/bench_c2.err; echo "config 2 rc=$?"; cat $OUT/bench_c2.json; tail -5 $OUT/bench_c2.err
python - > $OUT/csa.txt 2>&1 <<'PY'
import sys, time, numpy as np
sys.path.insert(0, '.')
import ctypes as C
from orb_slam2_b200 import _lib
lib = _lib.load()
import subprocess
for csa in (1, 0):
    lib.borb_debug_set_bow_csa(csa)
    # reuse bench_configs' config4 in-process
    sys.argv = ['x']
    import importlib, tools.bench_configs as bc
    r = bc.config4(2000, 12)
    print('csa', csa, 'all_ms', r['search_by_bow_all_ms'], 'counts_only_ms', r['search_by_bow_all_counts_only_ms'], 'top20_us', r['search_by_bow_top20_us'], 'query_us', r['kfdb_query_us'])
PY
cat $OUT/csa.txt | tail -4
