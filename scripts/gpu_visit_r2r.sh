#!/bin/bash
# Round 2, visit r: FP32 backward's weight-gradient GEMMs of all layers in one launch per chunk (fewer no-op launches in the guarded re-run).
set -u
OUT=gpurun_out/${1:-r2r}
mkdir -p $OUT
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -q -x -m gpu 2>&1 | tail -4 | tee $OUT/pytest_gpu.txt
echo "== backward timing"
for sh in cfg3 cfg2 cfg5 cfg4m; do timeout 300 python scripts/bwd_time.py $sh 20 2>&1 | tail -1 | tee -a $OUT/bwd_time.txt; done
UMNN_B200_PRECISION=fp32 timeout 300 python scripts/bwd_time.py cfg3 3 2>&1 | tail -1 | tee -a $OUT/bwd_time.txt
