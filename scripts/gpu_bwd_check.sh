#!/bin/bash
# Backward-focused GPU visit: gradient parity tests, then timings of the tensor-core backward per config.
set -u
TAG=${1:-bwdchk}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests -x -q -m gpu -k "backward or grad or training or flow or fused" 2>&1 | tail -6 | tee $OUT/pytest_bwd.txt
for c in cfg3 cfg2 cfg5 cfg4s; do timeout 300 python scripts/bwd_tc_bringup.py $c 2>&1 | tail -1; done | tee $OUT/bwd_times.txt
timeout 300 python scripts/train_step_bench.py auto 2>&1 | tail -1 | tee $OUT/train_step.txt
