#!/bin/bash
# Round 2, visit q: full GPU parity suite on the current build; per-pass launch lists of the backward for the narrow shapes (cfg5, cfg2).
set -u
OUT=gpurun_out/${1:-r2q}
mkdir -p $OUT
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -q -x -m gpu 2>&1 | tail -4 | tee $OUT/pytest_gpu.txt
for sh in cfg5 cfg2 cfg3; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches_bwd_$sh.csv \
      python scripts/bwd_tc_shape.py $sh > $OUT/launches_bwd_$sh.log 2>&1
  python - $OUT/launches_bwd_$sh.csv $sh <<'PY'
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
    print(f"  {sys.argv[2]:6s} {k:62s} n={len(v):3d} mean {sum(v)/len(v)/1e3:9.1f} us  total {sum(v)/1e3/3:9.1f} us/call")
PY
done | tee $OUT/passes.txt
