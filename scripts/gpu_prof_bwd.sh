#!/bin/bash
# ncu --set full captures of the three backward passes (pass F = forward kernel with operand emission).
set -u
TAG=${1:-profbwd}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for k in cc_forward_tc cc_dgrad_tc cc_wgrad_tc; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -o $OUT/prof_$k \
      python scripts/bwd_tc_only.py > $OUT/prof_$k.log 2>&1
done
python - <<'PY'
import torch, time
x = torch.empty(3 << 30, dtype=torch.uint8, device="cuda")
for _ in range(2): x.fill_(1)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(5): x.fill_(1)
e.record(); torch.cuda.synchronize()
print("write-only fill: %.1f GB/s" % (5 * x.numel() / (s.elapsed_time(e) * 1e-3) / 1e9))
y = torch.empty_like(x)
s.record()
for _ in range(5): y.copy_(x)
e.record(); torch.cuda.synchronize()
print("copy (read+write counted): %.1f GB/s" % (10 * x.numel() / (s.elapsed_time(e) * 1e-3) / 1e9))
s.record()
for _ in range(5): x.sum(dtype=torch.int64)
e.record(); torch.cuda.synchronize()
print("read-only sum: %.1f GB/s" % (5 * x.numel() / (s.elapsed_time(e) * 1e-3) / 1e9))
PY
ls -la $OUT
