#!/bin/bash
# GPU visit: full check of the round's state -- parity suite, smoke, bench lines for the five configurations and the
# reference arm, launch lists, ncu full captures (wide forward, narrow forward, the three backward passes), flow bench.
# Usage (under gpurun): bash scripts/gpu_visit_r1m.sh [tag]
set -u
TAG=${1:-r1m}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.txt 2>&1
nproc > $OUT/nproc.txt
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -30 | tee $OUT/pytest_gpu.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee $OUT/smoke.txt
echo "== forward bring-up"; timeout 900 python scripts/tc_bringup.py 2>&1 | grep -v "^---" | tee $OUT/bringup.txt | cut -c1-700
echo "== backward bring-up"; timeout 900 python scripts/bwd_tc_bringup.py 2>&1 | grep -v "^---" | tee $OUT/bwd_bringup.txt | cut -c1-400
echo "== bench cfg4"; timeout 900 python bench.py --steps 3 --warmup 3 2>&1 | tail -1 | tee $OUT/bench_cfg4.json
for wl in cfg1 cfg2 cfg3 cfg5; do
  echo "== bench $wl"; timeout 300 python bench.py --workload $wl --steps 20 --warmup 5 --no-cpu 2>&1 | tail -1 | tee $OUT/bench_$wl.json
done
echo "== bench cfg4 bf16x3"; timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu --precision bf16x3 2>&1 | tail -1 | tee $OUT/bench_cfg4_bf16x3.json
echo "== bench cfg4 fp32 kernel"; timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu --precision fp32 2>&1 | tail -1 | tee $OUT/bench_cfg4_fp32.json
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 | tee $OUT/bench_reference.json
echo "== flow bench"; timeout 900 python scripts/flow_bench.py 2>&1 | grep "^{" | tee $OUT/flow_bench.jsonl | cut -c1-300
echo "== train step"; timeout 600 python scripts/train_step_bench.py auto fp32 2>&1 | tail -3 | tee $OUT/train_step.txt
echo "== ncu launch list (cfg4, B=8192)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 1 --batch 8192 --no-cpu > $OUT/launches_bench.log 2>&1
echo "== ncu full capture of the forward kernel (cfg4, B=8192)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:cc_forward -s 2 -c 1 -o $OUT/prof_fwd \
    python bench.py --steps 1 --warmup 1 --batch 8192 --no-cpu > $OUT/prof_bench.log 2>&1
echo "== ncu full capture of the narrow forward kernel (cfg5)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:cc_forward -s 2 -c 1 -o $OUT/prof_fwd_cfg5 \
    python bench.py --workload cfg5 --steps 1 --warmup 1 --no-cpu > $OUT/prof_bench_cfg5.log 2>&1
echo "== ncu captures of the backward passes (cfg3)"
for k in cc_forward_tc cc_dgrad_tc cc_wgrad_tc; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -o $OUT/prof_bwd_$k \
      python scripts/bwd_tc_only.py > $OUT/prof_bwd_$k.log 2>&1
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $OUT/launches_bwd.csv \
    python scripts/bwd_tc_only.py > $OUT/launches_bwd.log 2>&1
ls -la $OUT
