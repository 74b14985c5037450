"""Global matching against a scribble (few hundred labelled reference pixels, cfg.TEST_MODE): per-kernel times of the two
tcgen05 engines (profile slots) and wall time per call."""
import ctypes, sys, time
sys.path.insert(0, ".")
import torch
import bench
from cvpr2020_manet_b200 import _lib
from cvpr2020_manet_b200.config import cfg
from cvpr2020_manet_b200.networks import IntVOS as api
L = _lib.lib()
cfg.TEST_MODE = True
ref, prev, cur, ref_lab, prev_lab = bench.synth_inputs(1000)
H, W, N = bench.H, bench.W, bench.N_IDS
scr = torch.full((H, W), -1, dtype=torch.int32)
for o in range(N):
    scr[10 + 15 * o, 20:150] = o
    scr[5 + 15 * o:25 + 15 * o, 30 + 25 * o] = o
r, q, lab = ref.cuda().permute(1, 2, 0), cur.cuda().permute(1, 2, 0), scr.cuda().unsqueeze(-1)
mem = torch.ones(H, W, N, 1, device="cuda")
def run(engine, cache, n=20):
    api.FORCE_EXACT3_ENGINE = engine == "exact3"
    L.manet_profile_enable(n + 2); L.manet_profile_reset()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(n):
        api.nearest_neighbor_features_per_object(r, q, lab, 1, torch.tensor(N - 1), normalize=True, memory_frame=mem, reference_cache=cache)
    torch.cuda.synchronize(); wall = (time.perf_counter() - t0) / n * 1e6
    res = {"wall_us": round(wall, 1)}
    for slot, name in ((5, "prepass"), (0, "gemm"), (3, "refine"), (4, "rescan"), (6, "exact3")):
        buf = (ctypes.c_float * (n + 2))(); k = ctypes.c_int(0)
        L.manet_profile_read(slot, buf, n + 2, ctypes.byref(k))
        v = [buf[i] for i in range(k.value)][2:]
        res[name] = round(sum(v) / len(v) * 1e3, 1) if v else None
    L.manet_profile_enable(0); api.FORCE_EXACT3_ENGINE = False
    return res
print("labelled reference pixels:", int((scr >= 0).sum()))
print("auto                   :", run("auto", None))
print("auto, cached           :", run("auto", api.ReferenceOperands()))
print("three-product          :", run("exact3", None))
