"""Correlation op: our kernels against the reference's OWN native extension, run side by side on the GPU.

oracle/_ref/correlation_cuda_ref.so is the reference's correlation_cuda.cc + correlation_cuda_kernel.cu compiled
for sm_100a by oracle/build_ref_correlation.py (test infrastructure; nothing of it is in the repository or on the
product path).  Same argument lists as the reference's pybind module (correlation_cuda.cc:169-172).
kernel_size = 1 only: for kernel_size > 1 the reference reads out of bounds."""
import importlib.util
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

SO = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "correlation_cuda_ref.so")


@pytest.fixture(scope="module")
def ref_mod():
    if not os.path.exists(SO):
        pytest.skip("oracle/_ref/correlation_cuda_ref.so not built (python oracle/build_ref_correlation.py where /root/reference is mounted)")
    spec = importlib.util.spec_from_file_location("correlation_cuda_ref", SO)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


CASES = [  # B, C, H, W, pad, kernel, max_disp, stride1, stride2
    (1, 100, 30, 38, 4, 1, 4, 1, 1),      # MANet's historical use: pad = md = d, k = 1 (._bak/networks_old/IntVOS.py:264)
    (1, 100, 60, 107, 12, 1, 12, 1, 1),   # half-resolution 480p, d = 12
    (2, 16, 21, 19, 6, 1, 6, 1, 2),
    (1, 8, 24, 32, 4, 1, 4, 2, 2),
    (2, 5, 9, 11, 0, 1, 0, 1, 1),
    (2, 16, 23, 45, 3, 1, 5, 1, 1),       # fast path (k = 1, strides 1, C % 4 == 0) with pad != max_displacement, ragged tiles
    (1, 64, 17, 70, 7, 1, 7, 1, 1),       # fast path, even number of 16-byte words per pixel (padded shared-memory pitch)
    (1, 128, 12, 33, 16, 1, 16, 1, 1),    # fast path at its limits: C = 128, max_displacement = 16
]


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.float64])
@pytest.mark.parametrize("case", CASES)
def test_forward_backward_match_the_reference_extension(ref_mod, case, dtype):
    from cvpr2020_manet_b200.correlation_package import correlation_cuda as ours
    B, C, H, W, pad, ks, md, s1, s2 = case
    gen = torch.Generator().manual_seed(3)
    x = torch.rand(B, C, H, W, generator=gen).cuda().to(dtype)
    y = torch.rand(B, C, H, W, generator=gen).cuda().to(dtype)

    def fwd(mod):
        r1, r2, out = x.new_empty(0), x.new_empty(0), x.new_empty(0)
        assert mod.forward(x, y, r1, r2, out, pad, ks, md, s1, s2, 1) == 1
        torch.cuda.synchronize()
        return r1, r2, out

    r1a, r2a, oa = fwd(ref_mod)
    r1b, r2b, ob = fwd(ours)
    assert oa.shape == ob.shape and r1a.shape == r1b.shape
    assert torch.equal(r1a, r1b) and torch.equal(r2a, r2b)                 # the zero-padded NHWC copies: bit-exact
    tol = {torch.float32: 2e-6, torch.float16: 2e-3, torch.float64: 1e-12}[dtype]
    assert float((oa.double() - ob.double()).abs().max()) <= tol * max(1.0, float(oa.double().abs().max()))

    if s1 != 1:
        return      # stride1 > 1: the reference's backward truncates its index range (correlation_cuda_kernel.cu:172-176)
    g = torch.rand(oa.shape, generator=gen).cuda().to(dtype)

    def bwd(mod, r1, r2):
        g1, g2 = x.new_empty(0), x.new_empty(0)
        assert mod.backward(x, y, r1, r2, g, g1, g2, pad, ks, md, s1, s2, 1) == 1
        torch.cuda.synchronize()
        return g1, g2

    g1a, g2a = bwd(ref_mod, r1a, r2a)
    g1b, g2b = bwd(ours, r1b, r2b)
    btol = {torch.float32: 1e-5, torch.float16: 5e-3, torch.float64: 1e-11}[dtype]
    for a, b in ((g1a, g1b), (g2a, g2b)):
        assert a.shape == b.shape
        assert float((a.double() - b.double()).abs().max()) <= btol * max(1.0, float(a.double().abs().max()))
