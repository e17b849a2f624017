#!/bin/bash
# Round 2, visit i: wait-loop variants re-measured with the uniform-datapath issuer (forward shapes + config-3 backward).
set -u
OUT=gpurun_out/r2i
mkdir -p $OUT
for v in lds_timer lds_plain lds_count lds_plain_hint lds_timer; do
  export UMNN_B200_LIB=$PWD/umnn_b200/variants/libumnn_b200_$v.so
  for wl in "cfg4 --batch 8192" "cfg3" "cfg5" "cfg2"; do
    name=$(echo $wl | cut -d' ' -f1)
    timeout 300 python bench.py --workload $wl --steps 50 --warmup 10 --no-cpu --no-train 2>&1 | grep "^{" | tail -1 > $OUT/bench_${v}_$name.json
    python - $OUT/bench_${v}_$name.json $v $name <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(f"{sys.argv[2]:16s} {sys.argv[3]:5s} {d['ms_per_step']:.4f} ms  parity {d['parity']['integral_max_rel_err_vs_oracle']:.2e}")
except Exception as e:
    print(sys.argv[2], sys.argv[3], "ERR", e)
PY
  done
  timeout 300 python scripts/bwd_time.py cfg3 10 2>&1 | tail -1 | sed "s/^/$v /"
done | tee $OUT/variants.txt
unset UMNN_B200_LIB
