"""Correlation op at MANet's use ([1,100,60,107] and [1,100,120,214], pad = max_displacement = 12, kernel 1, strides 1):
this package's kernels next to the REFERENCE's own kernels (oracle/_ref/correlation_cuda_ref.so, built from /root/reference
for sm_100a by oracle/build_ref_correlation.py), forward and backward, CUDA events.  Profiling script (not product code).
Usage: python scripts/corr_times.py [out.json]"""
import importlib.util
import json
import os
import sys

sys.path.insert(0, ".")
import torch  # noqa: E402

from cvpr2020_manet_b200.correlation_package import correlation_cuda as ours  # noqa: E402

SO = os.path.join("oracle", "_ref", "correlation_cuda_ref.so")
ref = None
if os.path.exists(SO):
    spec = importlib.util.spec_from_file_location("correlation_cuda_ref", SO)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)


def timed(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters * 1e3      # us


rows = []
for (H, W) in ((60, 107), (120, 214)):
    gen = torch.Generator().manual_seed(0)
    x = torch.rand(1, 100, H, W, generator=gen).cuda()
    y = torch.rand(1, 100, H, W, generator=gen).cuda()
    pad = md = 12
    rec = {"shape": [1, 100, H, W], "pad": pad, "max_displacement": md}
    for name, mod in (("ours", ours), ("reference_kernels", ref)):
        if mod is None:
            continue
        r1, r2, out = x.new_empty(0), x.new_empty(0), x.new_empty(0)
        mod.forward(x, y, r1, r2, out, pad, 1, md, 1, 1, 1)
        g = torch.rand_like(out)
        g1, g2 = x.new_empty(0), x.new_empty(0)
        rec[name] = {"forward_us": timed(lambda: mod.forward(x, y, r1, r2, out, pad, 1, md, 1, 1, 1)),
                     "backward_us": timed(lambda: mod.backward(x, y, r1, r2, g, g1, g2, pad, 1, md, 1, 1, 1))}
    if "reference_kernels" in rec:
        rec["speedup_forward"] = rec["reference_kernels"]["forward_us"] / rec["ours"]["forward_us"]
        rec["speedup_backward"] = rec["reference_kernels"]["backward_us"] / rec["ours"]["backward_us"]
    rows.append(rec)
    print(json.dumps(rec))
if len(sys.argv) > 1:
    json.dump({"what": "Correlation forward/backward, ours vs the reference's kernels compiled for sm_100a, same GPU, CUDA events, 20 iterations",
               "rows": rows}, open(sys.argv[1], "w"), indent=1)
