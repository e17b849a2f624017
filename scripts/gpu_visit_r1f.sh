#!/bin/bash
# GPU visit: narrow (2 CTAs/SM) kernel shape and FP16x3 operand split -- bring-up numbers, full parity suite both
# ways, bench lines.  Usage (under gpurun): bash scripts/gpu_visit_r1f.sh [tag]
set -u
TAG=${1:-r1f}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu.txt 2>&1
echo "== bring-up (per-case subprocesses)"; timeout 900 python scripts/tc_bringup.py 2>&1 | tee $OUT/bringup.txt | tail -40
echo "== pytest -m gpu (AUTO -> bf16x3)"; timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee $OUT/pytest_gpu.txt
echo "== pytest -m gpu (AUTO -> fp16x3)"; UMNN_B200_AUTO_TC=fp16x3 timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -25 | tee $OUT/pytest_gpu_fp16.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee $OUT/smoke.txt
for wl in cfg2 cfg5 cfg1 cfg3; do
  for nr in 1 0; do
    echo "== bench $wl narrow=$nr"; UMNN_B200_TC_NARROW=$nr timeout 300 python bench.py --workload $wl --steps 20 --warmup 5 --no-cpu 2>&1 | tail -1 | tee $OUT/bench_${wl}_narrow$nr.json
  done
done
echo "== bench cfg4 bf16x3"; timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu --precision bf16x3 2>&1 | tail -1 | tee $OUT/bench_cfg4_bf16x3.json
echo "== bench cfg4 fp16x3"; timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu --precision fp16x3 2>&1 | tail -1 | tee $OUT/bench_cfg4_fp16x3.json
ls -la $OUT
