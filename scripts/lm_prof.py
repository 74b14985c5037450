"""Minimal driver for profiling the local-matching kernels (480p, N=6, d=12)."""
import sys
import torch
sys.path.insert(0, ".")
from cvpr2020_manet_b200.networks import IntVOS as api  # noqa: E402

C, H, W, N, d = 100, 120, 214, 6, 12
torch.manual_seed(0)
p = torch.rand(C, H, W).cuda().permute(1, 2, 0)
q = torch.rand(C, H, W).cuda().permute(1, 2, 0)
lab = torch.randint(0, N, (H // 8 + 1, W // 8 + 1)).repeat_interleave(8, 0).repeat_interleave(8, 1)[:H, :W].int().cuda().unsqueeze(-1).contiguous()
ids = torch.arange(N).int().cuda()
api.FORCE_SIMT_LOCAL_ENGINE = len(sys.argv) > 1 and sys.argv[1] == "simt"
for _ in range(4):
    api.local_previous_frame_nearest_neighbor_features_per_object(p, q, lab, ids, d)
torch.cuda.synchronize()
