"""Quick GPU probe of the tcgen05 global-matching engine against the CUDA-core engine.
Usage: python scripts/umma_probe.py [H W Hr N]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cvpr2020_manet_b200 import _lib
from cvpr2020_manet_b200.networks import IntVOS as api

H, W, Hr, N = (int(x) for x in (sys.argv[1:5] if len(sys.argv) >= 5 else (24, 30, 24, 3)))
C = 100
gen = torch.Generator().manual_seed(0)
for dist in ("A", "B"):
    mk = (lambda *s: torch.rand(*s, generator=gen)) if dist == "A" else (lambda *s: 0.1 * torch.relu(torch.randn(*s, generator=gen)))
    ref = mk(C, Hr, W).cuda().permute(1, 2, 0)
    qry = mk(C, H, W).cuda().permute(1, 2, 0)
    lab = torch.randint(0, N, (Hr, W, 1), generator=gen).int().cuda()
    lab[lab == N - 1] = 0 if N > 2 else lab[lab == N - 1]
    api.FORCE_SIMT_ENGINE = True
    want, _ = api.nearest_neighbor_features_per_object(ref, qry, lab, 1, torch.tensor(N - 1))
    torch.cuda.synchronize()
    api.FORCE_SIMT_ENGINE = False
    t0 = time.time()
    got, _ = api.nearest_neighbor_features_per_object(ref, qry, lab, 1, torch.tensor(N - 1))
    torch.cuda.synchronize()
    dt = time.time() - t0
    absent = want == 1e20
    same_absent = bool(torch.equal(got == 1e20, absent))
    err = ((got - want).abs() / want.abs().clamp(min=1.0))[~absent]
    print(f"dist {dist}: M={H*W} R={Hr*W} N={N} absent_ok={same_absent} max_rel_err={float(err.max()):.3e} "
          f"mean={float(err.mean()):.3e} first call {dt*1e3:.1f} ms", flush=True)
    if float(err.max()) > 1e-4:
        bad = (((got - want).abs() / want.abs().clamp(min=1.0)) > 1e-4) & ~absent
        idx = bad.nonzero()[:8]
        print("  bad entries (b,y,x,o,_):", idx.tolist())
        print("  got ", got[bad][:8].tolist())
        print("  want", want[bad][:8].tolist())
        print("  bad fraction", float(bad.float().mean()))
