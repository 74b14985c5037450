"""Where the two global-matching engines cross: time both (forced) at several reference sizes, 480p query, N=6.
Usage: python scripts/gm_breakeven.py"""
import sys

sys.path.insert(0, ".")
import torch  # noqa: E402

import bench  # noqa: E402
from cvpr2020_manet_b200.networks import IntVOS as api  # noqa: E402

ref, prev, cur, ref_lab, prev_lab = bench.synth_inputs(1000)
q = cur.cuda().permute(1, 2, 0)
rfull = ref.cuda().permute(1, 2, 0).reshape(-1, 1, bench.C)
lfull = ref_lab.cuda().reshape(-1, 1, 1)
mem = None
for tiles in (40, 56, 64, 72, 88, 100):
    n = tiles * 256 - 6 * 128          # ~tiles after per-object padding
    r, l = rfull[:n], lfull[:n]
    res = {}
    for name, fr in (("exact3", False), ("fr", True)):
        api.FORCE_FR_ENGINE, api.FORCE_EXACT3_ENGINE = fr, not fr
        for _ in range(3):
            api.nearest_neighbor_features_per_object(r, q, l, 1, torch.tensor(bench.N_IDS - 1), normalize=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            api.nearest_neighbor_features_per_object(r, q, l, 1, torch.tensor(bench.N_IDS - 1), normalize=True)
        e1.record(); torch.cuda.synchronize()
        res[name] = e0.elapsed_time(e1) * 100.0
    print("R=%6d (~%3d tiles): exact3 %.1f us, filter-and-refine %.1f us" % (n, tiles, res["exact3"], res["fr"]))
api.FORCE_FR_ENGINE = api.FORCE_EXACT3_ENGINE = False
