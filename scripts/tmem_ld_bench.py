"""tcgen05.ld-only micro-benchmark (manet_microbench_tmem_ld): bytes per clock per SM that epilogue warps can drain
from tensor memory.  Usage: python scripts/tmem_ld_bench.py [out.json]   (writes JSON, prints a table)"""
import ctypes
import json
import sys

sys.path.insert(0, ".")
import torch  # noqa: E402  (creates the CUDA context the library runs in)

from cvpr2020_manet_b200 import _lib  # noqa: E402

torch.cuda.init()
torch.zeros(1, device="cuda")
L = _lib.lib()
MODES = {0: "ld.32x32b.x32, 2 in flight", 1: "ld.32x32b.x64, wait each", 2: "ld.x32 + global-matching epilogue maxima"}
rows = []
iters = 4000
for ctas in (1, 148):
    for mode in (0, 1, 2):
        for warps in (4, 8, 16):
            buf = (ctypes.c_longlong * ctas)()
            for rep in range(2):   # first call warms the instruction cache
                _lib.check(L.manet_microbench_tmem_ld(mode, iters, warps, ctas, buf, None), "tmem_ld bench")
            cyc = max(buf[i] for i in range(ctas))
            nbytes = iters * 128 * 512 * 4
            rows.append({"ctas": ctas, "mode": MODES[mode], "warps": warps, "cycles": int(cyc),
                         "bytes_per_clk_per_sm": nbytes / cyc,
                         "cycles_per_128x256_fp32_tile": 128 * 256 * 4 / (nbytes / cyc)})
            print(f"ctas {ctas:4d}  warps {warps:2d}  {MODES[mode]:45s} {nbytes / cyc:8.1f} B/clk/SM   "
                  f"{128 * 256 * 4 / (nbytes / cyc):7.0f} clk per 128x256 fp32 tile")
out = {"what": "tcgen05.ld-only drain of all 512 TMEM columns x 128 lanes, no MMA in flight (csrc/microbench.cu)",
       "iters": iters, "rows": rows, "best_bytes_per_clk_per_sm": max(r["bytes_per_clk_per_sm"] for r in rows)}
if len(sys.argv) > 1:
    json.dump(out, open(sys.argv[1], "w"), indent=1)
