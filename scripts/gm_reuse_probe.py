"""Pre-pass time of global matching with and without the reference cache (profile slot 5), 480p bench tensors."""
import ctypes, sys
sys.path.insert(0, ".")
import torch
import bench
from cvpr2020_manet_b200 import _lib
from cvpr2020_manet_b200.networks import IntVOS as api
L = _lib.lib()
ref, prev, cur, ref_lab, prev_lab = bench.synth_inputs(1000)
r, q, lab = ref.cuda().permute(1, 2, 0), cur.cuda().permute(1, 2, 0), ref_lab.cuda().unsqueeze(-1)
def run(cache, n=12):
    L.manet_profile_enable(n + 2); L.manet_profile_reset()
    for i in range(n):
        api.nearest_neighbor_features_per_object(r, q, lab, 1, torch.tensor(bench.N_IDS - 1), normalize=True, reference_cache=cache)
    torch.cuda.synchronize()
    res = {}
    for slot, name in ((5, "prepass"), (0, "filter"), (3, "refine"), (4, "rescan")):
        buf = (ctypes.c_float * (n + 2))(); k = ctypes.c_int(0)
        L.manet_profile_read(slot, buf, n + 2, ctypes.byref(k))
        v = [buf[i] for i in range(k.value)][2:]
        res[name] = round(sum(v) / len(v) * 1e3, 1)
    L.manet_profile_enable(0)
    return res
print("no cache :", run(None))
print("cache    :", run(api.ReferenceOperands()))
