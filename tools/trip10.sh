#!/bin/bash
TAG=${1:-t10}
OUT=gpurun_out/r02/$TAG
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1; echo "rc=$?" >> $OUT/pytest.log
tail -8 $OUT/pytest.log
timeout 300 python tools/fast_ablation.py > $OUT/fast_ablation.json 2> $OUT/fast_ablation.err; cat $OUT/fast_ablation.json; tail -3 $OUT/fast_ablation.err
for c in 1 2 4; do
  timeout 900 python bench.py --config $c --steps 20 --warmup 3 > $OUT/bench_c$c.json 2> $OUT/bench_c$c.err; echo "config $c rc=$?"; cat $OUT/bench_c$c.json; tail -3 $OUT/bench_c$c.err
done
