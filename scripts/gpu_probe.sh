#!/bin/bash
# Runs the tcgen05 hardware probe over a list of (mode, N, K) cases; each case in its own process
# under a timeout so a trap or hang cannot take the others down.
OUT=gpurun_out/${1:-probe}
mkdir -p $OUT
P=umnn_b200/csrc/probe/tc_probe
run() { echo "--- tc_probe $*"; timeout 60 $P "$@" 2>&1 | tail -12; echo "rc=$?"; }
{
run 0 64 16
run 0 208 64
run 0 208 208
run 1 16 16
run 1 208 64
run 1 208 208
run 3 256 64
run 3 208 64
run 2 256 64
run 2 192 64
run 2 224 64
run 2 208 64
run 2 208 208
run 2 224 208
} 2>&1 | tee $OUT/probe.txt
nvidia-smi --query-gpu=name,memory.used --format=csv | tee -a $OUT/probe.txt
