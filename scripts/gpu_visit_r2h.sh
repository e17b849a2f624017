#!/bin/bash
# Round 2, visit h (8 GPUs): the driver's SCALE command at N=8 -- strong-scaled forward + NCCL training legs -- and the NCCL bucket check.
set -u
OUT=gpurun_out/r2h
mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv > $OUT/gpu.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
echo "== bench N=8"; timeout 600 $TR --master-port 29544 bench.py --gpus 8 --steps 5 --warmup 3 2>&1 | grep "^{" | tail -1 | tee $OUT/bench_cfg4_8gpu.json | cut -c1-3000
echo "== ddp check N=8"; timeout 300 $TR --master-port 29546 scripts/ddp_check.py 2>&1 | grep "ddp_check\|DDP_CHECK" | tee $OUT/ddp_check.txt | cut -c1-700
ls $OUT
