"""Staged check of the tcgen05 local-matching engine against the CUDA-core engine (same process,
same inputs): (1) the window-distance volume, (2) the final per-object map.  Prints, does not assert."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from cvpr2020_manet_b200.networks import IntVOS as api  # noqa: E402


def run(H, W, C, N, d, seed=0, mode="noisy"):
    gen = torch.Generator().manual_seed(seed)
    if mode == "noisy":
        prev = 0.1 * torch.relu(torch.randn(C, H, W, generator=gen))
        cur = prev + 0.02 * torch.randn(C, H, W, generator=gen)
    elif mode == "rand":
        prev = torch.rand(C, H, W, generator=gen)
        cur = torch.rand(C, H, W, generator=gen)
    else:  # self
        prev = torch.rand(C, H, W, generator=gen)
        cur = prev.clone()
    lab = torch.randint(0, N, (H // 5 + 1, W // 5 + 1), generator=gen).repeat_interleave(5, 0).repeat_interleave(5, 1)[:H, :W].int()
    ids = torch.arange(N).int().cuda()
    p = prev.cuda().permute(1, 2, 0)
    q = cur.cuda().permute(1, 2, 0)
    lab = lab.cuda().unsqueeze(-1)
    res = {}
    for eng in ("simt", "tcgen05"):
        api.FORCE_SIMT_LOCAL_ENGINE = eng == "simt"
        vol = api.local_pairwise_distances2(q, p, d)
        out = api.local_previous_frame_nearest_neighbor_features_per_object(p, q, lab, ids, d)
        torch.cuda.synchronize()
        res[eng] = (vol.cpu().numpy(), out.cpu().numpy())
    v0, o0 = res["simt"]
    v1, o1 = res["tcgen05"]
    ev = np.abs(v0 - v1)
    eo = np.abs(o0 - o1)
    print(f"[{mode}] H={H} W={W} C={C} N={N} d={d}: volume max err {ev.max():.3e} (mean {ev.mean():.2e}), "
          f"map max err {eo.max():.3e} (mean {eo.mean():.2e}); nan vol {np.isnan(v1).sum()} map {np.isnan(o1).sum()}", flush=True)
    if ev.max() > 1e-5:
        idx = np.argwhere(ev > 1e-5)
        print("   volume mismatches:", len(idx), "first:", idx[:6].tolist(),
              [(float(v0[tuple(i)]), float(v1[tuple(i)])) for i in idx[:6]])
        D2 = 2 * d + 1
        ys, xs, ls = idx[:, 0], idx[:, 1], idx[:, 2]
        print("   by Y%16:", np.bincount(ys % 16, minlength=16).tolist())
        print("   by dy:", np.bincount(ls // D2, minlength=D2).tolist())
        print("   by dx:", np.bincount(ls % D2, minlength=D2).tolist())
    if eo.max() > 1e-5:
        idx = np.argwhere(eo > 1e-5)
        print("   map mismatches:", len(idx), "first:", idx[:6].tolist(),
              [(float(o0[tuple(i)]), float(o1[tuple(i)])) for i in idx[:6]])
        print("   by Y%14:", np.bincount(idx[:, 1] % 14, minlength=14).tolist())
        print("   by X%30:", np.bincount(idx[:, 2] % 30, minlength=30).tolist())
        print("   by obj:", np.bincount(idx[:, 3], minlength=N).tolist())


if __name__ == "__main__":
    t0 = time.time()
    shapes = [(20, 22, 100, 3, 5), (60, 106, 100, 6, 12), (120, 214, 100, 6, 12), (120, 214, 100, 6, 9), (33, 47, 64, 4, 7)]
    for (H, W, C, N, d) in shapes:
        for mode in ("noisy", "rand", "self"):
            try:
                run(H, W, C, N, d, mode=mode)
            except Exception as e:  # noqa: BLE001
                print("FAILED", (H, W, C, N, d, mode), repr(e), flush=True)
                raise
    # timing
    C, H, W, N, d = 100, 120, 214, 6, 12
    p = torch.rand(C, H, W).cuda().permute(1, 2, 0)
    q = torch.rand(C, H, W).cuda().permute(1, 2, 0)
    lab = torch.randint(0, N, (H, W, 1)).int().cuda()
    ids = torch.arange(N).int().cuda()
    for eng in ("simt", "tcgen05"):
        api.FORCE_SIMT_LOCAL_ENGINE = eng == "simt"
        for _ in range(3):
            api.local_previous_frame_nearest_neighbor_features_per_object(p, q, lab, ids, d)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            api.local_previous_frame_nearest_neighbor_features_per_object(p, q, lab, ids, d)
        e1.record()
        torch.cuda.synchronize()
        print(f"{eng}: {e0.elapsed_time(e1) / 20 * 1000:.1f} us per local match (incl. python)")
    print("done in", time.time() - t0, "s")
