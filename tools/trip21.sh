#!/bin/bash
TAG=${1:-t21}
OUT=gpurun_out/r02/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_match.py tests/test_gpu_frame.py tests/test_gpu_adapters.py tests/test_gpu_kfdb.py tests/test_gpu_voc_real.py -m gpu -x -q > $OUT/pytest.log 2>&1; echo "rc=$?" >> $OUT/pytest.log
tail -4 $OUT/pytest.log
timeout 900 python bench.py --config 2 --steps 20 --warmup 3 --no-cpu-baseline > $OUT/bench_c2.json 2> $OUT/bench_c2.err; echo "config 2 rc=$?"; python -c "
import json;d=json.load(open('$OUT/bench_c2.json'));print(d['value'],d['e2e']['value'],d.get('matcher_latency'))"; tail -3 $OUT/bench_c2.err
timeout 900 python bench.py --config 4 --steps 20 --warmup 3 --no-cpu-baseline > $OUT/bench_c4.json 2> $OUT/bench_c4.err; echo "config 4 rc=$?"; python -c "
import json;d=json.load(open('$OUT/bench_c4.json'));print(d['value'],d['e2e']['value'],d['roofline']['mean_launch_ms'])"; tail -3 $OUT/bench_c4.err
timeout 600 python tools/bench_configs.py > $OUT/configs.json 2> $OUT/configs.err; cat $OUT/configs.json; tail -3 $OUT/configs.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 2130 -c 300 --csv --log-file $OUT/launches_c4.csv python bench.py --config 4 --steps 2 --warmup 1 --no-cpu-baseline --no-parity > $OUT/ncu_c4.log 2>&1
echo "ncu c4 rc=$? lines=$(wc -l < $OUT/launches_c4.csv)"
