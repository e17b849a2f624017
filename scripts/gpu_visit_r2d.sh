#!/bin/bash
# Round 2, visit d (2 GPUs): NCCL training-step check, strong-scaled bench at N=2, reference arm under torchrun.
set -u
OUT=gpurun_out/${1:-r2d}
mkdir -p $OUT
nvidia-smi --query-gpu=index,name,clocks.sm,power.draw --format=csv > $OUT/gpu.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
echo "== ddp check"; timeout 600 $TR --master-port 29533 scripts/ddp_check.py 2>&1 | grep -v "^W\|^\[W\|OMP_NUM" | tail -6 | tee $OUT/ddp_check.txt | cut -c1-900
echo "== bench N=2"; timeout 900 $TR --master-port 29534 bench.py --gpus 2 --steps 5 --warmup 3 2>&1 | grep "^{" | tail -1 | tee $OUT/bench_cfg4_2gpu.json | cut -c1-3000
echo "== bench reference arm N=2"; timeout 600 $TR --master-port 29535 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>&1 | grep "^{" | tail -1 | tee $OUT/bench_reference_2gpu.json | cut -c1-400
echo "== bench N=1 (same box)"; timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu 2>&1 | grep "^{" | tail -1 | tee $OUT/bench_cfg4_1gpu.json | cut -c1-1500
ls -la $OUT
