#!/bin/bash
# Round 2, visit l: work-conserving MMA issue order (segment 1 advances while segment 0 waits for its operand) against the
# strict segment order, same box: forward shapes, config-3 backward, per-pass launch list.
set -u
OUT=gpurun_out/${1:-r2l}
mkdir -p $OUT
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -q -x -m gpu 2>&1 | tail -4 | tee $OUT/pytest_gpu.txt
for v in default issue_inorder default; do
  if [ $v = default ]; then unset UMNN_B200_LIB; else export UMNN_B200_LIB=$PWD/umnn_b200/variants/libumnn_b200_$v.so; fi
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
  for sh in cfg3 cfg5; do timeout 300 python scripts/bwd_time.py $sh 10 2>&1 | tail -1 | sed "s/^/$v /"; done
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/launches_bwd_$v.csv \
      python scripts/bwd_tc_only.py > $OUT/launches_bwd.log 2>&1
  python - $OUT/launches_bwd_$v.csv $v <<'PY'
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1], errors="replace")))
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
    if "tc_kernel" in k: print(f"  {sys.argv[2]:14s} {k:58s} n={len(v):3d} mean {sum(v)/len(v)/1e3:9.1f} us")
PY
done | tee $OUT/variants.txt
unset UMNN_B200_LIB
