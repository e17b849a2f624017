#!/bin/bash
# Round 2, visit n: pass W of chunk c on a side stream / its own SMs while passes F and D of chunk c + 1 run.
set -u
OUT=gpurun_out/${1:-r2n}
mkdir -p $OUT
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -q -x -m gpu 2>&1 | tail -4 | tee $OUT/pytest_gpu.txt
echo "== backward timing"
for sh in cfg4m cfg3x4 cfg3; do
  UMNN_B200_BWD_OVERLAP=0 timeout 300 python scripts/bwd_time.py $sh 2>&1 | tail -1 | tee -a $OUT/bwd_time.txt
  for w in 16 24 32 40; do
    UMNN_B200_BWD_OVERLAP=1 UMNN_B200_BWD_WSMS=$w timeout 300 python scripts/bwd_time.py $sh 2>&1 | tail -1 | tee -a $OUT/bwd_time.txt
  done
done
timeout 300 python scripts/bwd_time.py cfg4m 2>&1 | tail -1 | tee -a $OUT/bwd_time.txt
timeout 300 python scripts/bwd_time.py cfg3 2>&1 | tail -1 | tee -a $OUT/bwd_time.txt
echo "== flow bench"; timeout 900 python scripts/flow_bench.py bsds --no-torch 2>&1 | grep "^{" | tee $OUT/flow_bench.jsonl | cut -c1-400
