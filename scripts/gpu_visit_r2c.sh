#!/bin/bash
# Round 2, visit c: multi-block W stages (half-panel layout), invert with per-dimension conditioner outputs, small-call latency.
set -u
OUT=gpurun_out/r2c
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.txt 2>&1
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -25 | tee $OUT/pytest_gpu.txt
echo "== backward timing, W stage size A/B"
for kbs in 1 2 3 4; do UMNN_B200_WGRAD_KBS=$kbs timeout 300 python scripts/bwd_time.py cfg3 2>&1 | tail -1 | tee -a $OUT/bwd_time.txt; done
for sh in cfg3 cfg2 cfg5 cfg4s cfg4m; do timeout 300 python scripts/bwd_time.py $sh 2>&1 | tail -1 | tee -a $OUT/bwd_time.txt; done
UMNN_B200_BWD_PANELS=hilo timeout 300 python scripts/bwd_time.py cfg3 2>&1 | tail -1 | tee -a $OUT/bwd_time.txt
echo "== launch list backward cfg3 (default)"
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
for k, v in agg.items(): print(f"  {k:62s} n={len(v):3d} mean {sum(v)/len(v)/1e3:9.1f} us  max {max(v)/1e3:9.1f} us  total {sum(v)/1e6:8.2f} ms")
PY
echo "== grad flip analysis (default panels)"; timeout 600 python scripts/grad_flip_analysis.py cfg3s cfg4s 2>&1 | grep -v "panels=a_hilo" | tee $OUT/grad_flip.txt | cut -c1-330
echo "== flow bench (kernel route)"; timeout 900 python scripts/flow_bench.py power mnist --no-torch 2>&1 | grep "^{" | tee $OUT/flow_bench.jsonl | cut -c1-400
echo "== flow bench mnist, invert graph"; UMNN_B200_INVERT_GRAPH=1 timeout 900 python scripts/flow_bench.py mnist --no-torch 2>&1 | grep "^{" | tee $OUT/flow_bench_graph.jsonl | cut -c1-400
echo "== bench cfg1"; timeout 300 python bench.py --workload cfg1 --steps 200 --warmup 20 --no-cpu --no-train 2>&1 | tail -1 | tee $OUT/bench_cfg1.json | cut -c1-700
echo "== bench cfg5"; timeout 300 python bench.py --workload cfg5 --steps 50 --warmup 10 --no-cpu --no-train 2>&1 | tail -1 | tee $OUT/bench_cfg5.json | cut -c1-300
ls -la $OUT
