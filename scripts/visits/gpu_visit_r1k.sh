#!/bin/bash
# GPU visit: A/B of the compile-time variants (shared-space vs generic smem base, wait-loop styles) on three workloads.
set -u
TAG=${1:-r1k}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== occupancy test"; timeout 300 python -m pytest tests -q -m gpu -k "narrow_shape or overflow" 2>&1 | tail -5 | tee $OUT/pytest_subset.txt
for v in lds_count gen_count lds_timer gen_timer lds_plain gen_plain lds_plain_hint; do
  export UMNN_B200_LIB=$PWD/umnn_b200/variants/libumnn_b200_$v.so
  for wl in "cfg4 --batch 8192" "cfg3" "cfg5"; do
    name=$(echo $wl | cut -d' ' -f1)
    timeout 300 python bench.py --workload $wl --steps 30 --warmup 5 --no-cpu 2>&1 | tail -1 > $OUT/bench_${v}_$name.json
    python - $OUT/bench_${v}_$name.json $v $name <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(f"{sys.argv[2]:16s} {sys.argv[3]:5s} {d['ms_per_step']:.4f} ms  e2e {d['e2e']['ms_per_step']:.4f} ms  parity {d['parity']['integral_max_rel_err_vs_oracle']:.2e}")
except Exception as e:
    print(sys.argv[2], sys.argv[3], "ERR", e)
PY
  done
done | tee $OUT/variants.txt
unset UMNN_B200_LIB
