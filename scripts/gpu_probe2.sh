#!/bin/bash
OUT=gpurun_out/${1:-probe2}
mkdir -p $OUT
P=umnn_b200/csrc/probe/tc_probe
run() { echo "--- tc_probe $*"; timeout 60 $P "$@" 2>&1 | tail -10; echo "rc=$?"; }
{
run 4 64 16
run 4 208 64
run 4 32 128
run 5 64 16
run 5 208 64
run 5 208 128
run 5 32 64
} 2>&1 | tee $OUT/probe.txt
