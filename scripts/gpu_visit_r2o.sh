#!/bin/bash
# Round 2, visit o: prepared launches (umnn_b200.prepare_integral), host-side launch memo (func attributes / occupancy / SM count).
set -u
OUT=gpurun_out/${1:-r2o}
mkdir -p $OUT
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -q -x -m gpu 2>&1 | tail -4 | tee $OUT/pytest_gpu.txt
for wl in cfg1 cfg2 cfg5 cfg3; do
  timeout 300 python bench.py --workload $wl --steps 500 --warmup 50 --no-cpu --no-train 2>&1 | grep "^{" | tail -1 > $OUT/bench_$wl.json
  python -c "
import json
d=json.loads(open('$OUT/bench_$wl.json').read())
print('$wl', 'ms/step', round(d['ms_per_step'],4), 'prepared', round(d['aux']['prepared_call']['ms_per_step'],4), 'e2e ms', round(d['e2e']['ms_per_step'],4), 'err', d['parity']['integral_max_rel_err_vs_oracle'])"
done
echo "== flow bench"; timeout 900 python scripts/flow_bench.py toy mnist --no-torch 2>&1 | grep "^{" | tee $OUT/flow_bench.jsonl | cut -c1-400
