#!/bin/bash
# Round 2, visit x: one N segment per MMA layer (UMNN_B200_TC_SEGMENTS=1) against the default two, narrow and wide shapes.
set -u
OUT=gpurun_out/${1:-r2x}
mkdir -p $OUT
for seg in 2 1 2 1; do
  if [ $seg = 1 ]; then export UMNN_B200_TC_SEGMENTS=1; else unset UMNN_B200_TC_SEGMENTS; fi
  for wl in cfg5 cfg2 cfg3 cfg1; do
    timeout 300 python bench.py --workload $wl --steps 200 --warmup 20 --no-cpu --no-train 2>&1 | grep "^{" | tail -1 > $OUT/bench_seg${seg}_$wl.json
    python - $OUT/bench_seg${seg}_$wl.json $seg $wl <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(f"segments={sys.argv[2]} {sys.argv[3]:5s} {d['ms_per_step']:.4f} ms  parity {d['parity']['integral_max_rel_err_vs_oracle']:.2e}")
except Exception as e:
    print(sys.argv[2], sys.argv[3], "ERR", e)
PY
  done
  for sh in cfg5 cfg2; do timeout 300 python scripts/bwd_time.py $sh 10 2>&1 | tail -1 | cut -c1-100 | sed "s/^/segments=$seg /"; done
done | tee $OUT/segments.txt
