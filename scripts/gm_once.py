"""A few global matches on bench.py's tensors (480p, N=6) -- the workload for ncu captures of the global-matching kernels -- and
the engine's diagnostics.  Usage: python scripts/gm_once.py [n_calls]"""
import ctypes
import sys

sys.path.insert(0, ".")
import torch  # noqa: E402

import bench  # noqa: E402
from cvpr2020_manet_b200 import _lib  # noqa: E402
from cvpr2020_manet_b200._device import stream_ptr, workspace  # noqa: E402
from cvpr2020_manet_b200.networks import IntVOS as api  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
ref, prev, cur, ref_lab, prev_lab = bench.synth_inputs(1000)
r, q, lab = ref.cuda().permute(1, 2, 0), cur.cuda().permute(1, 2, 0), ref_lab.cuda().unsqueeze(-1)
mem = torch.ones(bench.H, bench.W, bench.N_IDS, 1, device="cuda")
for i in range(n):
    out, _ = api.nearest_neighbor_features_per_object(r, q, lab, 1, torch.tensor(bench.N_IDS - 1), normalize=True, memory_frame=mem)
torch.cuda.synchronize()
L = _lib.lib()
dev = q.device
ws = workspace(dev, L.manet_global_match_workspace_bytes(bench.M_PIX, bench.M_PIX, bench.C, bench.N_IDS, 1), "global")
st = (ctypes.c_int32 * 4)()
_lib.check(L.manet_global_match_stats(ws.data_ptr(), st, stream_ptr(dev)), "stats")
print("global match stats: tiles %d segments %d rescan entries %d bias folded %d; checksum %.6f" % (st[0], st[1], st[2], st[3], float(out.sum())))
