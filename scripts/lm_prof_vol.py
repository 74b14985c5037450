"""Times the window-volume API (VOL mode of the tcgen05 kernel + upsample) vs the CUDA-core engine."""
import sys
import torch
sys.path.insert(0, ".")
from cvpr2020_manet_b200.networks import IntVOS as api  # noqa: E402

C, H, W, d = 100, 120, 214, 12
torch.manual_seed(0)
p = torch.rand(C, H, W).cuda().permute(1, 2, 0)
q = torch.rand(C, H, W).cuda().permute(1, 2, 0)
for eng in ("simt", "tcgen05"):
    api.FORCE_SIMT_LOCAL_ENGINE = eng == "simt"
    for _ in range(3):
        api.local_pairwise_distances2(q, p, d)
    torch.cuda.synchronize()
