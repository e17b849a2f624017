#!/bin/bash
# GPU visit: packed fp32 (FFMA2/FMUL2/FADD2) epilogue A/B, invert graph replay for many-dimensional flows, parity suite.
set -u
TAG=${1:-r1q}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -12 | tee $OUT/pytest_gpu.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $OUT/smoke.txt
for rep in 1 2; do
for v in f32x2_off f32x2_on; do
  export UMNN_B200_LIB=$PWD/umnn_b200/variants/libumnn_b200_$v.so
  for wl in "cfg4 --batch 8192" "cfg3" "cfg5" "cfg2"; do
    name=$(echo $wl | cut -d' ' -f1)
    timeout 300 python bench.py --workload $wl --steps 30 --warmup 5 --no-cpu 2>&1 | tail -1 > $OUT/bench_${v}_${name}_$rep.json
    python - $OUT/bench_${v}_${name}_$rep.json $v $name <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(f"{sys.argv[2]:10s} {sys.argv[3]:5s} {d['ms_per_step']:.4f} ms  e2e {d['e2e']['ms_per_step']:.4f} ms  parity {d['parity']['integral_max_rel_err_vs_oracle']:.2e}")
except Exception as e:
    print(sys.argv[2], sys.argv[3], "ERR", e)
PY
  done
done
done | tee $OUT/variants.txt
unset UMNN_B200_LIB
echo "== flow bench mnist/bsds (graph)"; timeout 900 python scripts/flow_bench.py mnist bsds --no-torch 2>&1 | grep "^{" | tee $OUT/flow_bench_graph.jsonl | cut -c1-330
echo "== flow bench mnist/bsds (no graph)"; UMNN_B200_INVERT_GRAPH=0 timeout 900 python scripts/flow_bench.py mnist bsds --no-torch 2>&1 | grep "^{" | tee $OUT/flow_bench_nograph.jsonl | cut -c1-330
echo "== bench cfg4 full"; timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu 2>&1 | tail -1 | tee $OUT/bench_cfg4.json | cut -c1-200
