"""Experiment builds of the library for gm_fr_kernel (not shipped): lib/libmanet_b200_<tag>.so with -DFR_TRACE (in-kernel cycle
trace), -DFR_EXP_NOCOMPUTE (accumulator loads only), -DFR_EXP_NOLOAD (arithmetic only).  scripts/gpu_r02_exp.sh runs them."""
import os
import subprocess
import sys

sys.path.insert(0, ".")
from cvpr2020_manet_b200 import build as b  # noqa: E402

b.build()
nvcc = b.nvcc_path()
VARIANTS = (("trace", "-DFR_TRACE"), ("nocompute", "-DFR_EXP_NOCOMPUTE"), ("noload", "-DFR_EXP_NOLOAD"))
if len(sys.argv) > 1:          # e.g. python scripts/build_fr_variants.py rw8=-DFR_REFINE_WARPS_N=8 rw6=-DFR_REFINE_WARPS_N=6
    VARIANTS = tuple(a.split("=", 1) for a in sys.argv[1:])
for tag, flag in VARIANTS:
    o = os.path.join(b.PKG, f"_obj_{tag}_gm.o")
    r = subprocess.run([nvcc] + b.NVCC_FLAGS + [flag, "-c", os.path.join(b.CSRC, "global_match_umma.cu"), "-o", o], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    objs = [os.path.join(b.OBJDIR, s.replace(".cu", ".o")) for s in b.SOURCES if s != "global_match_umma.cu"] + [o]
    r = subprocess.run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", os.path.join(b.LIBDIR, f"libmanet_b200_{tag}.so")] + objs,
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    print("built", tag)
