#!/bin/bash
# compute-sanitizer (memcheck, racecheck, synccheck) over one small local + global match through the public API
mkdir -p gpurun_out
cat > /tmp/san.py <<'PY'
import sys, torch
sys.path.insert(0, ".")
from cvpr2020_manet_b200.networks import IntVOS as api
C, H, W, N, d = 100, 34, 46, 4, 5
g = torch.Generator().manual_seed(0)
p = (0.1 * torch.relu(torch.randn(C, H, W, generator=g))).cuda().permute(1, 2, 0)
q = (0.1 * torch.relu(torch.randn(C, H, W, generator=g))).cuda().permute(1, 2, 0)
lab = torch.randint(0, N, (H, W, 1), generator=g).int().cuda()
ids = torch.arange(N).int().cuda()
for _ in range(2):
    api.local_previous_frame_nearest_neighbor_features_per_object(p, q, lab, ids, d)
    api.nearest_neighbor_features_per_object(p, q, lab, 1, N - 1)
torch.cuda.synchronize()
print("done")
PY
for tool in memcheck racecheck synccheck; do
  timeout -s KILL 500 compute-sanitizer --tool $tool --kernel-regex kns=manet python /tmp/san.py > gpurun_out/sanitize_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|done|Error|hazard" gpurun_out/sanitize_$tool.log | head -8
done
