#!/bin/bash
# compute-sanitizer passes over small launches of every kernel (memcheck everywhere, racecheck on the FFMA kernels)
OUT=gpurun_out/${1:-sanitize}
mkdir -p $OUT
for tool in ${SANITIZE_TOOLS:-memcheck racecheck}; do
  for which in ${SANITIZE_CASES:-fp32 bf16x3 fp16x3}; do
    echo "=== $tool $which"
    timeout 600 compute-sanitizer --tool $tool --error-exitcode 9 python scripts/sanitize_case.py $which 2>&1 | grep -vE "^$" | tail -12 | tee $OUT/${tool}_${which}.txt
    echo "rc=$?"
  done
done
