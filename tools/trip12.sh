#!/bin/bash
TAG=${1:-t12}
OUT=gpurun_out/r02/$TAG
mkdir -p $OUT
timeout 300 python tools/fast_ablation.py > $OUT/fast_ablation.json 2> $OUT/fast_ablation.err; cat $OUT/fast_ablation.json; tail -3 $OUT/fast_ablation.err
timeout 900 python bench.py --config 1 --steps 20 --warmup 3 --no-cpu-baseline > $OUT/bench_c1.json 2> $OUT/bench_c1.err; echo "config 1 rc=$?"; cat $OUT/bench_c1.json | python -c "import sys,json; d=json.load(sys.stdin); print(d['value'], d['e2e']['value'], d['stage_ms_per_pass'])"
