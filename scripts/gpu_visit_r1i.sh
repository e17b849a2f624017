#!/bin/bash
# GPU visit: mbarrier suspend hints, narrow shape with the full shared-memory carveout (2 CTAs/SM), staged host entry.
# Usage (under gpurun): bash scripts/gpu_visit_r1i.sh [tag]
set -u
TAG=${1:-r1i}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu.txt 2>&1
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -30 | tee $OUT/pytest_gpu.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee $OUT/smoke.txt
echo "== forward bring-up"; timeout 900 python scripts/tc_bringup.py 2>&1 | grep -v "^---" | tee $OUT/bringup.txt | cut -c1-600
echo "== backward bring-up"; timeout 900 python scripts/bwd_tc_bringup.py 2>&1 | grep -v "^---" | tee $OUT/bwd_bringup.txt | cut -c1-400
echo "== bench cfg4"; timeout 900 python bench.py --steps 3 --warmup 3 2>&1 | tail -1 | tee $OUT/bench_cfg4.json
for wl in cfg1 cfg2 cfg3 cfg5; do
  echo "== bench $wl"; timeout 300 python bench.py --workload $wl --steps 20 --warmup 5 --no-cpu 2>&1 | tail -1 | tee $OUT/bench_$wl.json
done
echo "== ncu launch list (cfg4, B=8192)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 1 --batch 8192 --no-cpu > $OUT/launches_bench.log 2>&1
echo "== ncu full capture of the forward kernel (cfg4, B=8192)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:cc_forward -s 2 -c 1 -o $OUT/prof_fwd \
    python bench.py --steps 1 --warmup 1 --batch 8192 --no-cpu > $OUT/prof_bench.log 2>&1
echo "== ncu full capture of the narrow forward kernel (cfg5)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:cc_forward -s 2 -c 1 -o $OUT/prof_fwd_cfg5 \
    python bench.py --workload cfg5 --steps 1 --warmup 1 --no-cpu > $OUT/prof_bench_cfg5.log 2>&1
ls -la $OUT
