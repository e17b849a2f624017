#!/bin/bash
# Round 2, visit a: parity suite after the FP32 re-run / hi-only panels, gradient flip analysis, backward timing per panel mode.
set -u
OUT=gpurun_out/r2a
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.txt 2>&1
nproc > $OUT/nproc.txt
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -40 | tee $OUT/pytest_gpu.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee $OUT/smoke.txt
echo "== grad flip analysis"; timeout 900 python scripts/grad_flip_analysis.py 2>&1 | tee $OUT/grad_flip_analysis.txt | cut -c1-400
for pm in hi a_hilo hilo; do
  echo "== backward bring-up panels=$pm"
  UMNN_B200_BWD_PANELS=$pm timeout 900 python scripts/bwd_tc_bringup.py cfg3 2>&1 | tail -2 | tee $OUT/bwd_bringup_$pm.txt | cut -c1-600
  UMNN_B200_BWD_PANELS=$pm timeout 900 python scripts/bwd_tc_bringup.py cfg4s 2>&1 | tail -2 | tee -a $OUT/bwd_bringup_$pm.txt | cut -c1-600
  UMNN_B200_BWD_PANELS=$pm timeout 900 python scripts/bwd_tc_bringup.py cfg5 2>&1 | tail -2 | tee -a $OUT/bwd_bringup_$pm.txt | cut -c1-600
done
echo "== bench cfg4"; timeout 900 python bench.py --steps 3 --warmup 3 2>&1 | tail -1 | tee $OUT/bench_cfg4.json | cut -c1-1500
echo "== bench cfg1"; timeout 300 python bench.py --workload cfg1 --steps 50 --warmup 10 --no-cpu 2>&1 | tail -1 | tee $OUT/bench_cfg1.json | cut -c1-600
ls -la $OUT
