"""Per-kernel GPU durations (CUPTI via torch.profiler) of back-to-back local matches and full session steps.
Usage: python scripts/kernel_times.py [simt]"""
import sys
import torch
from torch.profiler import profile, ProfilerActivity
sys.path.insert(0, ".")
from cvpr2020_manet_b200.networks import IntVOS as api  # noqa: E402

C, H, W, N, d = 100, 120, 214, 6, 12
torch.manual_seed(0)
p = (0.1 * torch.relu(torch.randn(C, H, W))).cuda().permute(1, 2, 0)
q = (p.permute(2, 0, 1) + 0.02 * torch.randn(C, H, W).cuda()).permute(1, 2, 0)
r = (0.1 * torch.relu(torch.randn(C, H, W))).cuda().permute(1, 2, 0)
lab = torch.randint(0, N, (H // 8 + 1, W // 8 + 1)).repeat_interleave(8, 0).repeat_interleave(8, 1)[:H, :W].int().cuda().unsqueeze(-1).contiguous()
ids = torch.arange(N).int().cuda()
api.FORCE_SIMT_LOCAL_ENGINE = len(sys.argv) > 1 and sys.argv[1] == "simt"


def step():
    api.local_previous_frame_nearest_neighbor_features_per_object(p, q, lab, ids, d)
    api.nearest_neighbor_features_per_object(r, q, lab, 1, N - 1)


for _ in range(5):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(20):
        step()
    torch.cuda.synchronize()
rows = [(e.key, e.device_time_total / max(e.count, 1), e.count) for e in prof.key_averages() if e.device_time_total > 0]
for k, t, n in sorted(rows, key=lambda x: -x[1]):
    print(f"{t:9.2f} us  x{n:4d}  {k[:110]}")
