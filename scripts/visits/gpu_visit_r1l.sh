#!/bin/bash
# GPU visit: A/B of explicit nanosleep back-off in the mbarrier wait loops (forward on three workloads + cfg3 backward).
set -u
TAG=${1:-r1l}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== narrow tests"; timeout 300 python -m pytest tests -q -m gpu -k "narrow_shape" 2>&1 | tail -5 | tee $OUT/pytest_subset.txt
for v in lds_timer lds_sleep20 lds_sleep50 lds_sleep100 lds_sleep200 lds_sleep50_nohint; do
  export UMNN_B200_LIB=$PWD/umnn_b200/variants/libumnn_b200_$v.so
  for wl in "cfg4 --batch 8192" "cfg3" "cfg5" "cfg2"; do
    name=$(echo $wl | cut -d' ' -f1)
    timeout 300 python bench.py --workload $wl --steps 30 --warmup 5 --no-cpu 2>&1 | tail -1 > $OUT/bench_${v}_$name.json
    python - $OUT/bench_${v}_$name.json $v $name <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(f"{sys.argv[2]:18s} {sys.argv[3]:5s} {d['ms_per_step']:.4f} ms  e2e {d['e2e']['ms_per_step']:.4f} ms  parity {d['parity']['integral_max_rel_err_vs_oracle']:.2e}")
except Exception as e:
    print(sys.argv[2], sys.argv[3], "ERR", e)
PY
  done
  echo "$v bwd $(timeout 300 python scripts/bwd_tc_bringup.py cfg3 2>&1 | tail -1 | cut -c1-120)"
done | tee $OUT/variants.txt
unset UMNN_B200_LIB
