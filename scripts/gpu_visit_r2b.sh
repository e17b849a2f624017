#!/bin/bash
# Round 2, visit b: full parity suite, new bench line (reference arms, training legs), per-pass timing of the backward
# (launch lists for hi-only and hi+lo panels) and ncu full captures of the three passes with hi-only panels.
set -u
OUT=gpurun_out/r2b
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.txt 2>&1
nproc > $OUT/nproc.txt
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -40 | tee $OUT/pytest_gpu.txt
echo "== bench cfg4 (default line)"; timeout 1200 python bench.py 2>&1 | tail -1 | tee $OUT/bench_cfg4.json | cut -c1-3000
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 | tee $OUT/bench_reference.json | cut -c1-800
for pm in hi hilo; do
  echo "== launch list backward cfg3 panels=$pm"
  UMNN_B200_BWD_PANELS=$pm timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $OUT/launches_bwd_$pm.csv \
      python scripts/bwd_tc_only.py > $OUT/launches_bwd_$pm.log 2>&1
  python - <<PY
import csv, collections
rows = list(csv.reader(open("$OUT/launches_bwd_$pm.csv", errors="replace")))
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
H = rows[hdr]; kn, mv = H.index("Kernel Name"), H.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) <= mv: continue
    name = r[kn].split("(")[0][:60]
    try: v = float(r[mv].replace(",", ""))
    except ValueError: continue
    agg.setdefault(name, []).append(v)
for k, v in agg.items(): print(f"  {k:62s} n={len(v):3d} mean {sum(v)/len(v)/1e3:9.1f} us  max {max(v)/1e3:9.1f} us  total {sum(v)/1e6:8.2f} ms")
PY
done
echo "== ncu captures of the backward passes (cfg3, hi-only panels)"
for k in cc_forward_tc cc_dgrad_tc cc_wgrad_tc; do
  UMNN_B200_BWD_PANELS=hi timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -o $OUT/prof_bwd_$k \
      python scripts/bwd_tc_only.py > $OUT/prof_bwd_$k.log 2>&1
done
ls -la $OUT
