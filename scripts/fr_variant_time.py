"""Time gm_fr_kernel (library profile slot 0) on bench.py's tensors.  Used with experiment builds of the library
(-DFR_EXP_NOCOMPUTE / -DFR_EXP_NOLOAD / -DFR_TRACE) copied over lib/libmanet_b200.so on the GPU box."""
import ctypes
import sys

sys.path.insert(0, ".")
import torch  # noqa: E402

import bench  # noqa: E402
from cvpr2020_manet_b200 import _lib  # noqa: E402
from cvpr2020_manet_b200.networks import IntVOS as api  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 12
ref, prev, cur, ref_lab, prev_lab = bench.synth_inputs(1000)
r, q, lab = ref.cuda().permute(1, 2, 0), cur.cuda().permute(1, 2, 0), ref_lab.cuda().unsqueeze(-1)
mem = torch.ones(bench.H, bench.W, bench.N_IDS, 1, device="cuda")
L = _lib.lib()
api.FORCE_FR_ENGINE = True
for i in range(3):
    api.nearest_neighbor_features_per_object(r, q, lab, 1, torch.tensor(bench.N_IDS - 1), normalize=True, memory_frame=mem)
torch.cuda.synchronize()
L.manet_profile_enable(n + 4)
L.manet_profile_reset()
for i in range(n):
    api.nearest_neighbor_features_per_object(r, q, lab, 1, torch.tensor(bench.N_IDS - 1), normalize=True, memory_frame=mem)
torch.cuda.synchronize()
for slot, name in ((5, "prepass"), (0, "gm_fr"), (3, "refine"), (4, "rescan")):
    buf = (ctypes.c_float * (n + 4))()
    cnt = ctypes.c_int(0)
    _lib.check(L.manet_profile_read(slot, buf, n + 4, ctypes.byref(cnt)), "read")
    v = sorted(buf[i] for i in range(cnt.value))
    print(name, "median_us %.1f min_us %.1f n=%d" % (v[len(v) // 2] * 1e3, v[0] * 1e3, cnt.value))
