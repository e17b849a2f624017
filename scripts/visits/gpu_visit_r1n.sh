#!/bin/bash
# GPU visit: same-box comparison of the generic-LD and LDS builds, isolated under ncu (launch list) and back to back.
set -u
TAG=${1:-r1n}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for rep in 1 2; do
for v in gen_timer lds_timer; do
  export UMNN_B200_LIB=$PWD/umnn_b200/variants/libumnn_b200_$v.so
  timeout 600 ncu --metrics gpu__time_duration.sum,sm__cycles_elapsed.avg.per_second --clock-control none -c 40 --csv --log-file $OUT/launches_${v}_$rep.csv \
      python bench.py --steps 2 --warmup 1 --batch 8192 --no-cpu > $OUT/launches_${v}_$rep.log 2>&1
  python - $OUT/launches_${v}_$rep.csv $v <<'PY'
import csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[0].isdigit()]
t = [float(r[14]) for r in rows if "cc_forward_tc_kernel<1, 0, 0, 0>" in r[4] and r[12] == "gpu__time_duration.sum"]
c = [float(r[14]) for r in rows if "cc_forward_tc_kernel<1, 0, 0, 0>" in r[4] and "per_second" in r[12]]
print(sys.argv[2], "ncu isolated: B=8192 launches", [round(x / 1e6, 3) for x in t if x > 1e7][:6], "clock GHz", [round(x, 3) for x in c[:3]])
PY
  timeout 300 python bench.py --workload cfg4 --batch 8192 --steps 30 --warmup 5 --no-cpu 2>&1 | tail -1 > $OUT/bench_${v}_$rep.json
  python - $OUT/bench_${v}_$rep.json $v <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[2], "bench back-to-back: %.3f ms/step" % d["ms_per_step"], d["clocks"])
PY
done
done | tee $OUT/ab.txt
