#!/bin/bash
# Round 2, visit e: uniform-register MMA issuers (forward, dgrad, wgrad) -- parity, forward / backward timings for every config.
set -u
OUT=gpurun_out/r2e
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.txt 2>&1
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -15 | tee $OUT/pytest_gpu.txt
for wl in cfg1 cfg2 cfg3 cfg5; do
  echo "== bench $wl"; timeout 300 python bench.py --workload $wl --steps 100 --warmup 10 --no-cpu --no-train 2>&1 | grep "^{" | tail -1 | tee $OUT/bench_$wl.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['config']['workload'][:40], 'ms/step', round(d['ms_per_step'],4), 'G evals/s', round(d['value']/1e9,3), 'e2e ms', round(d['e2e']['ms_per_step'],4), 'frac', round(d['roofline']['frac'],3), 'issued', round(d['roofline']['issued_tensor_frac'],3), 'err', d['parity']['integral_max_rel_err_vs_oracle'])"
done
echo "== bench cfg4"; timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-train 2>&1 | grep "^{" | tail -1 | tee $OUT/bench_cfg4.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['config']['workload'][:40], 'ms/step', round(d['ms_per_step'],3), 'G evals/s', round(d['value']/1e9,3), 'e2e ms', round(d['e2e']['ms_per_step'],3), 'frac', round(d['roofline']['frac'],3), 'issued', round(d['roofline']['issued_tensor_frac'],3), 'clocks', d['clocks'])"
echo "== backward timing"
for sh in cfg3 cfg2 cfg5 cfg4s cfg4m; do timeout 300 python scripts/bwd_time.py $sh 2>&1 | tail -1 | tee -a $OUT/bwd_time.txt; done
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
echo "== flow bench"; timeout 900 python scripts/flow_bench.py power bsds mnist --no-torch 2>&1 | grep "^{" | tee $OUT/flow_bench.jsonl | cut -c1-400
ls $OUT
