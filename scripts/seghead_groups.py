"""Does running the head per object group (smaller working set: more of it stays in the 126 MB L2) pay?"""
import sys
import torch
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
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def run(group):
    outs = []
    for g0 in range(0, N, group):
        outs.append(head.forward_parts(cur, gmap[:, :, :, g0:g0 + group].contiguous(), lmap[:, :, :, g0:g0 + group].contiguous(), prev,
                                       ids[g0:g0 + group]))
    return torch.cat(outs, 0)


ref = run(N)
for group in (6, 3, 2, 1):
    for _ in range(2):
        out = run(group)
    assert torch.equal(out, ref) or float((out - ref).abs().max()) < 1e-4, group
    ts = []
    for cold in (True, False):
        torch.cuda.synchronize()
        tot = 0.0
        for i in range(8):
            if cold:
                flush.fill_(i)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            run(group)
            e1.record()
            torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
        ts.append(tot / 8 * 1000)
    print(f"group of {group}: {ts[0]:.1f} us L2-cold, {ts[1]:.1f} us back to back")
