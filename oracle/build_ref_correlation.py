"""Build the reference's OWN native Correlation extension (correlation_package/correlation_cuda.cc +
correlation_cuda_kernel.cu, /root/reference) for sm_100a into oracle/_ref/correlation_cuda_ref.so.

TEST INFRASTRUCTURE ONLY (oracle side): the product never loads this file.  It exists so that the GPU parity
tests can run the reference's kernels themselves next to ours on identical inputs (tests/test_gpu_reference_corr.py).

Nothing of the reference is copied into the repository: the three source files are read where they lie, a
scratch copy under a temporary directory receives the one mechanical edit modern PyTorch needs
(`Tensor.type()` -> `Tensor.scalar_type()` inside AT_DISPATCH_*, 7 places in correlation_cuda_kernel.cu,
SURVEY.md section 8c), nvcc cross-compiles it (no GPU needed), and only the resulting .so lands in
oracle/_ref/ (git-ignored; it travels to the GPU box with the snapshot).

    python oracle/build_ref_correlation.py
"""
from __future__ import annotations

import os
import shutil
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(os.environ.get("MANET_REFERENCE_ROOT", "/root/reference"), "correlation_package")
OUT_DIR = os.path.join(HERE, "_ref")
NAME = "correlation_cuda_ref"
FILES = ["correlation_cuda.cc", "correlation_cuda_kernel.cu", "correlation_cuda_kernel.cuh"]


def reference_available() -> bool:
    return all(os.path.isfile(os.path.join(REF_DIR, f)) for f in FILES)


def built_path() -> str:
    return os.path.join(OUT_DIR, NAME + ".so")


def build(verbose: bool = False) -> str:
    if not reference_available():
        raise RuntimeError(f"reference sources not found under {REF_DIR}")
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0")
    from torch.utils.cpp_extension import load
    os.makedirs(OUT_DIR, exist_ok=True)
    with tempfile.TemporaryDirectory(prefix="manet_ref_corr_") as tmp:
        for f in FILES:
            text = open(os.path.join(REF_DIR, f)).read()
            if f.endswith(".cu"):
                text = text.replace(".type()", ".scalar_type()")
            open(os.path.join(tmp, f), "w").write(text)
        build_dir = os.path.join(tmp, "build")
        os.makedirs(build_dir)
        load(name=NAME, sources=[os.path.join(tmp, "correlation_cuda.cc"), os.path.join(tmp, "correlation_cuda_kernel.cu")],
             extra_cuda_cflags=["-gencode", "arch=compute_100a,code=sm_100a", "--expt-relaxed-constexpr", "-w"],
             extra_cflags=["-w"], build_directory=build_dir, verbose=verbose, is_python_module=False)
        shutil.copyfile(os.path.join(build_dir, NAME + ".so"), built_path())
    return built_path()


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv))
