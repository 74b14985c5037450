"""Per-kernel GPU durations (CUPTI via torch.profiler) of the DynamicSegHead forward at 480p, 5 objects.
Usage: python scripts/seghead_times.py"""
import sys
import torch
from torch.profiler import profile, ProfilerActivity
sys.path.insert(0, ".")
from cvpr2020_manet_b200.networks.seghead import DynamicSegHead  # noqa: E402

C, H, W, N = 100, 120, 214, 6
torch.manual_seed(0)
head = DynamicSegHead().cuda().eval()
cur = (0.1 * torch.relu(torch.randn(C, H, W))).cuda()
gmap = torch.rand(1, H, W, N, 1).cuda()
lmap = torch.rand(1, H, W, N, 1).cuda()
prev = torch.randint(0, N, (H // 8 + 1, W // 8 + 1)).repeat_interleave(8, 0).repeat_interleave(8, 1)[:H, :W].int().cuda()
ids = torch.arange(N).int().cuda()


def step():
    return head.forward_parts(cur, gmap, lmap, prev, ids)


for _ in range(3):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    step()
e1.record()
torch.cuda.synchronize()
print(f"seghead forward: {e0.elapsed_time(e1) / 10 * 1000:.1f} us per frame (events, back to back)")
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(10):
        step()
    torch.cuda.synchronize()
rows = [(e.key, e.device_time_total / max(e.count, 1), e.count) for e in prof.key_averages() if e.device_time_total > 0]
for k, t, n in sorted(rows, key=lambda x: -x[1]):
    print(f"{t:9.2f} us  x{n:4d}  {k[:110]}")
