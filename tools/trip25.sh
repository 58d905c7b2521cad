#!/bin/bash
# 2-GPU trip: NCCL entry points of the C ABI + torchrun bench of configs 1 and 4
TAG=${1:-t25}
OUT=gpurun_out/r02/$TAG
mkdir -p $OUT
N=${2:-2}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/nccl_cabi_check.py > $OUT/nccl_cabi.log 2>&1; echo "nccl rc=$?"; grep -v Warning $OUT/nccl_cabi.log | tail -3
for c in 1 4 2; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$c bench.py --gpus $N --config $c --steps 20 --warmup 3 > $OUT/bench_c${c}_n$N.json 2> $OUT/bench_c${c}_n$N.err; echo "config $c rc=$?"
  python -c "
import json
for l in open('$OUT/bench_c${c}_n$N.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['n_gpus'], d['value'], d['e2e']['value'], d.get('cpu_baseline'))"
  tail -2 $OUT/bench_c${c}_n$N.err
done
