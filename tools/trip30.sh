#!/bin/bash
TAG=${1:-t30}; N=${2:-4}
OUT=gpurun_out/r02/$TAG
mkdir -p $OUT
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 20 --warmup 3 > $OUT/bench_c1_n$N.json 2> $OUT/bench_c1_n$N.err; echo "config 1 rc=$?"
python -c "
import json
for l in open('$OUT/bench_c1_n$N.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['n_gpus'], d['value'], d['e2e']['value'], d.get('nccl'), d['clocks'])"
tail -3 $OUT/bench_c1_n$N.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 bench.py --impl reference --gpus $N --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c1-300
