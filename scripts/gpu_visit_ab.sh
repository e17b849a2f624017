#!/bin/bash
# Same-box A/B of the product build against umnn_b200/variants/libumnn_b200_<name>.so (forward shapes, backward, flow timings):
#   bash scripts/gpu_visit_ab.sh <out-tag> <variant-name>
set -u
OUT=gpurun_out/${1:-ab}
VAR=${2:-prev}
mkdir -p $OUT
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -q -x -m gpu 2>&1 | tail -3 | tee $OUT/pytest_gpu.txt
for v in default $VAR default $VAR; do
  if [ $v = default ]; then unset UMNN_B200_LIB; else export UMNN_B200_LIB=$PWD/umnn_b200/variants/libumnn_b200_$v.so; fi
  for wl in "cfg4 --batch 8192" "cfg3" "cfg5" "cfg2" "cfg1"; do
    name=$(echo $wl | cut -d' ' -f1)
    timeout 300 python bench.py --workload $wl --steps 100 --warmup 10 --no-cpu --no-train 2>&1 | grep "^{" | tail -1 > $OUT/bench_${v}_$name.json
    python - $OUT/bench_${v}_$name.json $v $name <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(f"{sys.argv[2]:10s} {sys.argv[3]:5s} {d['ms_per_step']:.4f} ms  prepared {d['aux']['prepared_call']['ms_per_step']:.4f}  parity {d['parity']['integral_max_rel_err_vs_oracle']:.2e}")
except Exception as e:
    print(sys.argv[2], sys.argv[3], "ERR", e)
PY
  done
  for sh in cfg3 cfg5 cfg2; do timeout 300 python scripts/bwd_time.py $sh 10 2>&1 | tail -1 | sed "s/^/$v /"; done
done | tee $OUT/variants.txt
unset UMNN_B200_LIB
