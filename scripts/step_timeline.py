"""Timeline of one propagation step of the session (two streams): where each profiled kernel group starts and ends relative to
the start of the global pre-pass.  Usage: python scripts/step_timeline.py [n_steps]"""
import ctypes
import sys

sys.path.insert(0, ".")
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from cvpr2020_manet_b200 import _lib  # noqa: E402
from cvpr2020_manet_b200.engine import MatchingSession  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
L = _lib.lib()
dev = torch.device("cuda:0")
sess = MatchingSession(bench.H, bench.W, bench.C, bench.N_IDS, bench.D_LOCAL, n_frames=104)
ref, prev, cur, ref_lab, prev_lab = bench.synth_inputs(1000)
for k, v in (("ref", ref), ("prev", prev), ("cur", cur), ("ref_labels", ref_lab), ("prev_labels", prev_lab)):
    sess.slots[0][k][:] = v
sess.upload()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for i in range(5):
    sess.step_device(1 + i, 1, 0, drop_unlabelled=False)
sess.sync()
L.manet_profile_enable(n + 4)
L.manet_profile_reset()
for i in range(n):
    flush.zero_()
    torch.cuda.synchronize()
    sess.step_device(1 + i % 100, 1, 0, drop_unlabelled=False)
sess.sync()
torch.cuda.synchronize()
names = {5: "global pre-pass", 0: "gm_fr", 3: "refine", 4: "rescan", 2: "local pre-pass", 1: "lm_umma"}
for slot in (5, 2, 1, 0, 3, 4):
    a = (ctypes.c_float * (n + 4))(); b = (ctypes.c_float * (n + 4))(); cnt = ctypes.c_int(0)
    _lib.check(L.manet_profile_read_span(slot, 5, a, b, n + 4, ctypes.byref(cnt)), "span")
    st = np.array(a[:cnt.value]) * 1e3; en = np.array(b[:cnt.value]) * 1e3
    print("%-16s start %7.1f us  end %7.1f us   (median of %d steps)" % (names[slot], np.median(st), np.median(en), cnt.value))
L.manet_profile_enable(0)
