#!/bin/bash
# final verification: full GPU suite, smoke, the three bench configs (defaults, with CPU baselines), reference arms
TAG=${1:-t34}
OUT=gpurun_out/r02/$TAG
mkdir -p $OUT
timeout 1800 python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1; echo "rc=$?" >> $OUT/pytest.log
tail -3 $OUT/pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
for c in 1 2 4; do
  timeout 900 python bench.py --config $c > $OUT/bench_c$c.json 2> $OUT/bench_c$c.err; echo "config $c rc=$?"
  python -c "
import json;d=json.load(open('$OUT/bench_c$c.json'));print(d['value'],d['e2e']['value'],d['roofline']['frac'],d['cpu_baseline']['value'] if d['cpu_baseline'] else None, d['cpu_baseline'].get('one_core_value') if d['cpu_baseline'] else None, d['ms_per_step']*d['steps'])"
  timeout 600 python bench.py --impl reference --config $c > $OUT/bench_ref_c$c.json 2>/dev/null; python -c "
import json;d=json.load(open('$OUT/bench_ref_c$c.json'));print('ref', d['value'], d['cpu_baseline']['cores'])"
done
