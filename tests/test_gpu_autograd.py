"""GPU parity of the gradients: the reference's matching functions are differentiable
torch graphs (train_stage1.py:126); here the CUDA backward kernels are compared with torch autograd through
the CPU oracle (a restatement of the reference's operation order, pinned bit-exact to the reference's outputs
by tests/test_oracle_golden.py) on the same seeded inputs.

Tolerance: gradients within 2e-4 of the largest gradient magnitude (fp32 summation order; the arg-min itself
is discrete, so the inputs are drawn from a continuous distribution where exact ties have probability zero).
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

GRAD_RTOL = 2e-4


@pytest.fixture(scope="module")
def api():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from cvpr2020_manet_b200.networks import IntVOS
    return IntVOS


def blobs(gen, h, w, n, cell=4):
    return torch.randint(0, n, (h // cell + 1, w // cell + 1), generator=gen).repeat_interleave(cell, 0).repeat_interleave(cell, 1)[:h, :w].int()


def rel_err(got, want):
    got, want = got.detach().cpu().double(), want.detach().cpu().double()
    return float((got - want).abs().max() / max(1e-12, float(want.abs().max())))


@pytest.mark.parametrize("permuted", [False, True])
@pytest.mark.parametrize("absent", [False, True])
def test_global_k1_gradients_match_oracle_autograd(api, permuted, absent):
    from oracle import manet_oracle as O
    gen = torch.Generator().manual_seed(11)
    hr, wr, h, w, C, n_obj = 20, 26, 18, 22, 32, 3
    ref = 0.3 * torch.randn(C, hr, wr, generator=gen)
    qry = 0.3 * torch.randn(C, h, w, generator=gen)
    lab = blobs(gen, hr, wr, n_obj + 1)
    if absent:
        lab[lab == 2] = 1                                   # object 2 has no reference pixel -> 1e20, zero gradient
    wts = torch.rand(1, h, w, n_obj + 1, 1, generator=gen)

    def run(fn, dev):
        r = ref.to(dev).clone().requires_grad_(True)
        q = qry.to(dev).clone().requires_grad_(True)
        rv, qv = (r.permute(1, 2, 0), q.permute(1, 2, 0)) if permuted else (r.permute(1, 2, 0).contiguous(), q.permute(1, 2, 0).contiguous())
        out, ids = fn(rv, qv, lab.to(dev).unsqueeze(-1), 1, torch.tensor(n_obj), 10)
        loss = (((torch.sigmoid(out) - 0.5) * 2) * wts.to(dev)).sum()      # the caller-side normalisation, IntVOS.py:611-612
        loss.backward()
        return out.detach().cpu(), r.grad.cpu(), q.grad.cpu()

    o_ref, gr_ref, gq_ref = run(O.global_match, "cpu")
    o_gpu, gr_gpu, gq_gpu = run(api.nearest_neighbor_features_per_object, "cuda")
    assert o_gpu.shape == o_ref.shape
    present = o_ref < 1e19
    assert torch.equal(o_gpu < 1e19, present)
    assert float(((o_gpu - o_ref).abs() / o_ref.abs().clamp(min=1.0))[present].max()) <= 2e-5
    assert rel_err(gq_gpu, gq_ref) <= GRAD_RTOL
    assert rel_err(gr_gpu, gr_ref) <= GRAD_RTOL
    assert float(gq_ref.abs().max()) > 0


@pytest.mark.parametrize("shape", [(24, 30, 16, 3, 3), (21, 27, 20, 4, 5)])
def test_local_gradients_match_oracle_autograd(api, shape):
    from oracle import manet_oracle as O
    H, W, C, n_ids, d = shape
    gen = torch.Generator().manual_seed(5)
    prev = 0.3 * torch.randn(C, H, W, generator=gen)
    cur = prev + 0.15 * torch.randn(C, H, W, generator=gen)
    lab = blobs(gen, H, W, n_ids)
    ids = torch.arange(n_ids, dtype=torch.int32)
    wts = torch.rand(1, H, W, n_ids, 1, generator=gen)

    def run(fn, dev):
        p = prev.to(dev).clone().requires_grad_(True)
        q = cur.to(dev).clone().requires_grad_(True)
        out = fn(p.permute(1, 2, 0), q.permute(1, 2, 0), lab.to(dev).unsqueeze(-1), ids.to(dev), d)
        (out * wts.to(dev)).sum().backward()
        return out.detach().cpu(), p.grad.cpu(), q.grad.cpu()

    o_ref, gp_ref, gq_ref = run(O.local_match, "cpu")
    o_gpu, gp_gpu, gq_gpu = run(api.local_previous_frame_nearest_neighbor_features_per_object, "cuda")
    assert float((o_gpu - o_ref).abs().max()) <= 1e-5
    assert float(gq_ref.abs().max()) > 0 and float(gp_ref.abs().max()) > 0
    assert rel_err(gq_gpu, gq_ref) <= GRAD_RTOL
    assert rel_err(gp_gpu, gp_ref) <= GRAD_RTOL


def test_no_grad_calls_keep_the_fast_engines(api):
    """requires_grad inputs under torch.no_grad() (inference, test.py:90) must not take the training path."""
    gen = torch.Generator().manual_seed(1)
    H, W, C, n = 24, 30, 100, 3
    p = (0.1 * torch.randn(C, H, W, generator=gen)).cuda().requires_grad_(True)
    q = (0.1 * torch.randn(C, H, W, generator=gen)).cuda().requires_grad_(True)
    lab = blobs(gen, H, W, n).cuda().unsqueeze(-1)
    with torch.no_grad():
        out, _ = api.nearest_neighbor_features_per_object(p.permute(1, 2, 0), q.permute(1, 2, 0), lab, 1, torch.tensor(n - 1), 10)
        loc = api.local_previous_frame_nearest_neighbor_features_per_object(p.permute(1, 2, 0), q.permute(1, 2, 0), lab,
                                                                          torch.arange(n).int().cuda(), 4)
    assert not out.requires_grad and not loc.requires_grad


def test_k_greater_than_one_with_grad_is_refused(api):
    gen = torch.Generator().manual_seed(2)
    r = torch.randn(8, 8, 16, generator=gen).cuda().requires_grad_(True)
    lab = torch.zeros(8, 8, 1, dtype=torch.int32).cuda()
    with pytest.raises(NotImplementedError):
        api.nearest_neighbor_features_per_object(r, r.detach(), lab, 3, torch.tensor(0), 10)


def test_gradients_match_reference_golden(api):
    """CUDA gradients against gradients recorded from the UNMODIFIED reference (tests/golden/make_golden.py)."""
    import os
    gdir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    g = np.load(os.path.join(gdir, "grad_global_k1.npz"))
    r = torch.from_numpy(g["ref_chw"]).cuda().requires_grad_(True)
    q = torch.from_numpy(g["query_chw"]).cuda().requires_grad_(True)
    out, _ = api.nearest_neighbor_features_per_object(r.permute(1, 2, 0), q.permute(1, 2, 0),
                                                      torch.from_numpy(g["labels"]).cuda().unsqueeze(-1), 1,
                                                      torch.tensor(int(g["n_obj"])), 5)
    (((torch.sigmoid(out) - 0.5) * 2) * torch.from_numpy(g["weights"]).cuda()).sum().backward()
    assert rel_err(r.grad, torch.from_numpy(g["grad_ref_chw"])) <= GRAD_RTOL
    assert rel_err(q.grad, torch.from_numpy(g["grad_query_chw"])) <= GRAD_RTOL

    g = np.load(os.path.join(gdir, "grad_local_d3.npz"))
    p = torch.from_numpy(g["prev_chw"]).cuda().requires_grad_(True)
    q = torch.from_numpy(g["cur_chw"]).cuda().requires_grad_(True)
    out = api.local_previous_frame_nearest_neighbor_features_per_object(
        p.permute(1, 2, 0), q.permute(1, 2, 0), torch.from_numpy(g["labels"]).cuda().unsqueeze(-1),
        torch.from_numpy(g["ids"]).cuda(), int(g["d"]))
    (out * torch.from_numpy(g["weights"]).cuda()).sum().backward()
    assert float((out.detach().cpu() - torch.from_numpy(g["out"])).abs().max()) <= 1e-5
    assert rel_err(p.grad, torch.from_numpy(g["grad_prev_chw"])) <= GRAD_RTOL
    assert rel_err(q.grad, torch.from_numpy(g["grad_query_chw"])) <= GRAD_RTOL
