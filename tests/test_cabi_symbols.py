"""CPU-side checks of the C-ABI boundary: the library builds/loads, exports every symbol the
header declares, and the product path refuses to run without its CUDA device (no fallback)."""
import ctypes
import os

import pytest
import torch

from cvpr2020_manet_b200 import _lib


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(_lib.LIB_PATH):
        from cvpr2020_manet_b200.build import build
        build()
    return _lib.lib()


def test_header_symbols_are_all_exported_and_bound(lib):
    declared = _lib.declared_symbols()
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/manet_b200.h but not exported"
    assert set(declared) == set(_lib.SIGNATURES), "ctypes table and header disagree"


def test_abi_version_and_error_string(lib):
    assert lib.manet_abi_version() == 1
    assert isinstance(lib.manet_last_error(), bytes)


def test_correlation_shape_math_matches_reference_formula(lib):
    # correlation_cuda.cc:31-34 ; MANet's use: pad = md = d, k = 1, strides 1 -> same H, W, (2d+1)^2 channels
    oc, oh, ow = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    assert lib.manet_correlation_output_shape(100, 120, 214, 9, 1, 9, 1, 1, oc, oh, ow) == 0
    assert (oc.value, oh.value, ow.value) == (361, 120, 214)
    assert lib.manet_correlation_output_shape(3, 64, 64, 20, 1, 20, 2, 2, oc, oh, ow) == 0   # FlowNet2-C setting
    assert (oc.value, oh.value, ow.value) == (441, 32, 32)
    assert lib.manet_correlation_output_shape(3, 8, 8, 0, 2, 0, 1, 1, oc, oh, ow) != 0       # even kernel rejected
    assert b"kernel_size" in lib.manet_last_error()


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_gpu_means_loud_failure_not_fallback(lib):
    assert lib.manet_check_device() == -3   # MANET_E_ARCH
    from cvpr2020_manet_b200.networks import IntVOS
    a = torch.rand(4, 5, 8)
    lab = torch.zeros(4, 5, 1, dtype=torch.int32)
    with pytest.raises(TypeError, match="no CPU fallback"):
        IntVOS.nearest_neighbor_features_per_object(a, a, lab, 1, torch.tensor(1))
    with pytest.raises(TypeError, match="no CPU fallback"):
        IntVOS.local_previous_frame_nearest_neighbor_features_per_object(a, a, lab, torch.arange(2).int(), 2)
    from cvpr2020_manet_b200.correlation_package.correlation import Correlation
    with pytest.raises(TypeError, match="no CPU fallback"):
        Correlation(2, 1, 2, 1, 1)(torch.rand(1, 3, 6, 6), torch.rand(1, 3, 6, 6))


def test_product_package_never_imports_the_oracle():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "cvpr2020_manet_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f


def test_reference_correlation_extension_loads_if_built():
    """oracle/_ref/correlation_cuda_ref.so (the reference's own extension, built by oracle/build_ref_correlation.py)
    imports and exposes the reference's pybind surface (correlation_cuda.cc:169-172).  No compute without a GPU."""
    import importlib.util
    import os

    import pytest
    so = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "correlation_cuda_ref.so")
    if not os.path.exists(so):
        pytest.skip("oracle/_ref not built here")
    import torch  # noqa: F401  (libtorch symbols)
    spec = importlib.util.spec_from_file_location("correlation_cuda_ref", so)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert callable(mod.forward) and callable(mod.backward)
