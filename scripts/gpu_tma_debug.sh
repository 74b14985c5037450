#!/bin/bash
mkdir -p gpurun_out
true
cat > /tmp/t.py <<'PY'
import sys, torch
sys.path.insert(0, ".")
from cvpr2020_manet_b200.networks.seghead import DynamicSegHead
torch.manual_seed(0)
head = DynamicSegHead().cuda().eval()
x = torch.randn(2, 103, 40, 70).cuda()
y = head(x); torch.cuda.synchronize(); print("ok", float(y.sum()))
PY
timeout -s KILL 300 compute-sanitizer --tool memcheck python /tmp/t.py > gpurun_out/tma_sanitizer.log 2>&1; echo "sanitizer rc=$?"; grep -E "Invalid|illegal|Illegal|at 0x|in .*sh_|ERROR SUMMARY|ok " gpurun_out/tma_sanitizer.log | head -20
