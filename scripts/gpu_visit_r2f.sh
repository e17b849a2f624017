#!/bin/bash
# Round 2, visit f: epoch-marked flag (no memset), umnn_invert_dimension, small-call latency, coalesced context gather.
set -u
OUT=gpurun_out/r2f
mkdir -p $OUT
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -15 | tee $OUT/pytest_gpu.txt
for wl in cfg1 cfg2 cfg5; do
  echo "== bench $wl"; timeout 300 python bench.py --workload $wl --steps 300 --warmup 20 --no-cpu --no-train 2>&1 | grep "^{" | tail -1 | tee $OUT/bench_$wl.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['config']['workload'][:40], 'ms/step', round(d['ms_per_step'],4), 'G evals/s', round(d['value']/1e9,3), 'e2e ms', round(d['e2e']['ms_per_step'],4), 'err', d['parity']['integral_max_rel_err_vs_oracle'])"
done
echo "== launch list cfg1"; timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $OUT/launches_cfg1.csv python bench.py --workload cfg1 --steps 5 --warmup 3 --no-cpu --no-train > $OUT/launches_cfg1.log 2>&1
python - <<PY
import csv
rows = list(csv.reader(open("$OUT/launches_cfg1.csv", errors="replace")))
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
H = rows[hdr]; kn, mv = H.index("Kernel Name"), H.index("Metric Value")
for r in rows[hdr + 1:][-12:]:
    if len(r) > mv: print("  ", r[kn].split("(")[0][:70], r[mv])
PY
echo "== flow bench"; timeout 900 python scripts/flow_bench.py toy power mnist --no-torch 2>&1 | grep "^{" | tee $OUT/flow_bench.jsonl | cut -c1-400
echo "== flow bench mnist, invert by rounds"; UMNN_B200_INVERT=rounds timeout 900 python scripts/flow_bench.py mnist --no-torch 2>&1 | grep "^{" | tee $OUT/flow_bench_rounds.jsonl | cut -c1-400
ls $OUT
