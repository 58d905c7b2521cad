#!/bin/bash
TAG=${1:-t19}
OUT=gpurun_out/r02/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_match.py tests/test_gpu_frame.py tests/test_gpu_adapters.py tests/test_gpu_kfdb.py -m gpu -x -q > $OUT/pytest.log 2>&1; echo "rc=$?" >> $OUT/pytest.log
tail -4 $OUT/pytest.log
timeout 900 python bench.py --config 2 --steps 20 --warmup 3 --no-cpu-baseline > $OUT/bench_c2.json 2> $OUT/bench_c2.err; echo "config 2 rc=$?"; python -c "
import json;d=json.load(open('$OUT/bench_c2.json'));print(d['value'],d['e2e']['value'],d['config'].get('matcher_latency'))"; tail -3 $OUT/bench_c2.err
timeout 600 python tools/bench_configs.py > $OUT/configs.json 2> $OUT/configs.err; cat $OUT/configs.json; tail -3 $OUT/configs.err
