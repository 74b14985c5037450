"""Error of the tcgen05 local-matching engine (forced, unguarded) against the exact CUDA-core engine as a function of the guard
statistic G = max |x - mu|^2 (both pooled frames): the measurement behind kLocalGuardG.  Post-BN-ReLU-like embeddings, scaled.
Usage: python scripts/lm_error_vs_g.py [out.json]"""
import json
import sys

sys.path.insert(0, ".")
import torch  # noqa: E402

from cvpr2020_manet_b200.networks import IntVOS as api  # noqa: E402

C, H, W, N, d = 100, 120, 214, 6, 12
rows = []
for kind in ("relu", "smooth"):
    for scale in (0.05, 0.1, 0.2, 0.4, 0.8, 1.6):
        gen = torch.Generator().manual_seed(int(scale * 1000) + len(kind))
        if kind == "relu":
            prev = scale * torch.relu(torch.randn(C, H, W, generator=gen))
            cur = prev + 0.1 * scale * torch.randn(C, H, W, generator=gen)
        else:   # spatially smooth field (bilinearly upsampled noise) + small change: nearby pixels are close, as in video
            base = torch.nn.functional.interpolate(torch.randn(1, C, H // 8, W // 8 + 1, generator=gen), size=(H, W), mode="bilinear")[0]
            prev = scale * torch.relu(base)
            cur = prev + 0.05 * scale * torch.randn(C, H, W, generator=gen)
        lab = torch.randint(0, N, (H // 8, W // 8 + 1), generator=gen).repeat_interleave(8, 0).repeat_interleave(8, 1)[:H, :W].int()
        p, q, l, ids = prev.cuda().permute(1, 2, 0), cur.cuda().permute(1, 2, 0), lab.cuda().unsqueeze(-1), torch.arange(N).int().cuda()
        api.FORCE_SIMT_LOCAL_ENGINE, api.FORCE_TENSOR_LOCAL_ENGINE = True, False
        exact = api.local_previous_frame_nearest_neighbor_features_per_object(p, q, l, ids, d)
        api.FORCE_SIMT_LOCAL_ENGINE, api.FORCE_TENSOR_LOCAL_ENGINE = False, True
        tens = api.local_previous_frame_nearest_neighbor_features_per_object(p, q, l, ids, d)
        api.FORCE_TENSOR_LOCAL_ENGINE = False
        st = api.local_match_guard_stats(H, W, C, N, d)
        err = float((tens - exact).abs().max())
        rows.append({"kind": kind, "scale": scale, "G": st["G"], "max_abs_err": err, "err_over_G": err / st["G"]})
        print(f"{kind:7s} scale {scale:4.2f}  G {st['G']:9.3f}  max |tensor - exact| {err:.3e}  err/G {err / st['G']:.2e}")
if len(sys.argv) > 1:
    json.dump({"what": "tcgen05 local matching (forced) vs exact CUDA-core engine, 480p, d=12, N=6", "rows": rows}, open(sys.argv[1], "w"), indent=1)
