#!/bin/bash
# Round 2, visit j: faster prep warps (register-blocked c_slot product, A_0 rows copied from the context rows, 32-bit
# row arithmetic), pair-major sign masks, parallel d_h reduction in pass D.
set -u
OUT=gpurun_out/${1:-r2j}
mkdir -p $OUT
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -q -x -m gpu 2>&1 | tail -8 | tee $OUT/pytest_gpu.txt
echo "== backward timing"
for sh in cfg3 cfg2 cfg5 cfg4m; do timeout 300 python scripts/bwd_time.py $sh 2>&1 | tail -1 | tee -a $OUT/bwd_time.txt; done
echo "== launch list backward cfg3"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $OUT/launches_bwd.csv \
    python scripts/bwd_tc_only.py > $OUT/launches_bwd.log 2>&1
python - <<PY
import csv, collections
rows = list(csv.reader(open("$OUT/launches_bwd.csv", errors="replace")))
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
H = rows[hdr]; kn, mv = H.index("Kernel Name"), H.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) <= mv: continue
    name = r[kn].split("(")[0][:60]
    try: v = float(r[mv].replace(",", ""))
    except ValueError: continue
    agg.setdefault(name, []).append(v)
for k, v in agg.items():
    if "tc_kernel" in k: print(f"  {k:62s} n={len(v):3d} mean {sum(v)/len(v)/1e3:9.1f} us")
PY
echo "== forward bench lines"
for wl in "cfg4 --batch 8192" cfg3 cfg5 cfg2 cfg1; do
  name=$(echo $wl | cut -d' ' -f1)
  timeout 300 python bench.py --workload $wl --steps 50 --warmup 10 --no-cpu --no-train 2>&1 | grep "^{" | tail -1 > $OUT/bench_$name.json
  python -c "
import json,sys
d=json.loads(open('$OUT/bench_$name.json').read())
print('$name', 'ms/step', round(d['ms_per_step'],4), 'e2e ms', round(d['e2e']['ms_per_step'],4), 'frac', round(d['roofline']['frac'],3), 'err', d['parity']['integral_max_rel_err_vs_oracle'])"
done
echo "== flow bench"; timeout 900 python scripts/flow_bench.py power bsds --no-torch 2>&1 | grep "^{" | tee $OUT/flow_bench.jsonl | cut -c1-400
