#!/bin/bash
# GPU visit (repeat after the mixed-operand fix): FP16x3 as the AUTO split (forward + backward re-evaluation, guarded bf16 re-runs), narrow kernel shape.
# Usage (under gpurun): bash scripts/gpu_visit_r1h.sh [tag]
set -u
TAG=${1:-r1h}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu.txt 2>&1
echo "== backward bring-up"; timeout 900 python scripts/bwd_tc_bringup.py 2>&1 | tee $OUT/bwd_bringup.txt | tail -30
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -30 | tee $OUT/pytest_gpu.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee $OUT/smoke.txt
echo "== bench cfg4"; timeout 900 python bench.py --steps 3 --warmup 3 2>&1 | tail -1 | tee $OUT/bench_cfg4.json
for wl in cfg1 cfg2 cfg3 cfg5; do
  echo "== bench $wl"; timeout 300 python bench.py --workload $wl --steps 20 --warmup 5 --no-cpu 2>&1 | tail -1 | tee $OUT/bench_$wl.json
done
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 | tee $OUT/bench_reference.json
echo "== flow bench"; timeout 900 python scripts/flow_bench.py 2>&1 | tail -12 | tee $OUT/flow_bench.txt
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
