#!/bin/bash
# One gpurun trip: GPU parity suite + measurements.  Outputs under gpurun_out/r02/<tag>/.
TAG=${1:-t1}
OUT=gpurun_out/r02/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest.log
tail -3 $OUT/pytest.log
timeout 300 python tools/fast_ablation.py > $OUT/fast_ablation.json 2> $OUT/fast_ablation.err; cat $OUT/fast_ablation.json
timeout 600 python tools/bench_configs.py > $OUT/configs.json 2> $OUT/configs.err; cat $OUT/configs.json
timeout 600 python bench.py --steps 200 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; cat $OUT/bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/match_launches.csv python tools/bench_configs.py --kfs 200 --reps 3 > $OUT/match_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bow_match_kernel -c 2 -o $OUT/bow_match_v0 python tools/bench_configs.py --kfs 2000 --reps 2 > $OUT/bow_ncu.log 2>&1
ls -la $OUT
