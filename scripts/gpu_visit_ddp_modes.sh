#!/bin/bash
# 2 GPUs: scripts/ddp_check.py with the default backward, with hi+lo operand panels and with the FP32 backward -- how much of the
# sharded-vs-single gradient difference is the hi-only panels' rounding (each run rounds ITS rows' operands to bf16).
set -u
OUT=gpurun_out/${1:-ddpm}
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
for mode in default hilo fp32; do
  case $mode in
    default) unset UMNN_B200_BWD_PANELS UMNN_B200_BACKWARD;;
    hilo) export UMNN_B200_BWD_PANELS=hilo; unset UMNN_B200_BACKWARD;;
    fp32) unset UMNN_B200_BWD_PANELS; export UMNN_B200_BACKWARD=fp32;;
  esac
  echo "== $mode"; timeout 600 $TR --master-port 2954$((RANDOM % 10)) scripts/ddp_check.py 2>&1 | grep "ddp_check\|DDP_CHECK" | sed "s/^/$mode /" | tee -a $OUT/ddp_modes.txt | cut -c1-700
done
