#!/bin/bash
# Same-box A/B of the backward: product build against umnn_b200/variants/libumnn_b200_<name>.so
#   bash scripts/gpu_visit_ab_bwd.sh <out-tag> <variant-name>
set -u
OUT=gpurun_out/${1:-abb}
VAR=${2:-prev}
mkdir -p $OUT
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -q -x -m gpu 2>&1 | tail -3 | tee $OUT/pytest_gpu.txt
for v in default $VAR default $VAR; do
  if [ $v = default ]; then unset UMNN_B200_LIB; else export UMNN_B200_LIB=$PWD/umnn_b200/variants/libumnn_b200_$v.so; fi
  for sh in cfg3 cfg5 cfg2 cfg4m; do timeout 300 python scripts/bwd_time.py $sh 10 2>&1 | tail -1 | cut -c1-110 | sed "s/^/$v /"; done
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
    if "tc_kernel" in k: print(f"  {sys.argv[2]:10s} {k:58s} n={len(v):3d} mean {sum(v)/len(v)/1e3:9.1f} us")
PY
done | tee $OUT/variants.txt
unset UMNN_B200_LIB
