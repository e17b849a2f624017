#!/bin/bash
# Round 2, visit ab: the rank-1 head panels (DZ_J, DZ_{J+1}) keep their lo part; gradient error of a flow per panel mode, per-layer
# error per cotangent structure, parity suite, backward timings.
set -u
OUT=gpurun_out/${1:-r2ab}
mkdir -p $OUT
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -q -x -m gpu 2>&1 | tail -3 | tee $OUT/pytest_gpu.txt
echo "== flow gradient vs FP32 backward"; for b in 500 4001 20000; do timeout 300 python scripts/panel_precision_check.py $b 2>&1 | grep "mode=" | cut -c1-120; done | tee $OUT/panel_precision.txt
echo "== per tensor (default panels)"; UMNN_B200_BWD_PANELS=hi_head timeout 300 python scripts/panel_precision_check.py 4001 --per-tensor 2>&1 | grep "parallel_nets" | cut -c1-170 | head -8 | tee -a $OUT/panel_precision.txt
echo "== per layer, one integrand call"; timeout 300 python scripts/panel_layer_error.py 4000 2>&1 | tail -7 | cut -c1-330 | tee $OUT/panel_layer_error.txt
echo "== backward timing"
for sh in cfg3 cfg2 cfg5 cfg4m; do timeout 300 python scripts/bwd_time.py $sh 10 2>&1 | tail -1 | cut -c1-100 | tee -a $OUT/bwd_time.txt; done
for sh in cfg3 cfg4m; do UMNN_B200_BWD_PANELS=hi timeout 300 python scripts/bwd_time.py $sh 10 2>&1 | tail -1 | cut -c1-100 | tee -a $OUT/bwd_time.txt; done
