"""The filter-and-refine global-matching engine (csrc/global_match_umma.cu: gm_fr_kernel + gm_refine_kernel +
gm_rescan_kernel): one tensor-core product of the fp16 hi parts filters candidates under a rigorous error bound, the
survivors are re-evaluated exactly in fp32.  These tests attack what is specific to it -- near-tied and exactly tied
references (every candidate path: second candidates, the rescan work list, its in-place overflow fallback), multi-tile
segments, extreme operand scales, the arg-min output -- against the fp32 CUDA-core engine, the three-product tensor engine
and brute force in float64.  Tolerances as everywhere: raw |a-b| <= 2e-5 max(1,|b|), sentinels bit-exact."""
import ctypes

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

RAW_RTOL = 2e-5


@pytest.fixture(scope="module")
def api():
    from cvpr2020_manet_b200 import _lib
    _lib.check(_lib.lib().manet_check_device(), "device check")
    from cvpr2020_manet_b200.networks import IntVOS
    return IntVOS


@pytest.fixture()
def option():
    """manet_set_option with automatic restore."""
    from cvpr2020_manet_b200 import _lib
    L = _lib.lib()
    saved = []

    def set_(name, value):
        saved.append((name, L.manet_set_option(name.encode(), int(value))))

    yield set_
    for name, old in reversed(saved):
        L.manet_set_option(name.encode(), old)


def rel_err(got, want):
    got, want = got.double().cpu(), want.double().cpu()
    return float(((got - want).abs() / want.abs().clamp_min(1.0)).max())


def brute_force(ref, qry, lab, n_obj):
    """float64 min_r |q - r|^2 per object and the lowest arg-min index; absent objects: 1e20 / -1."""
    d = torch.cdist(qry.double(), ref.double()) ** 2
    vals, idxs = [], []
    for o in range(n_obj):
        m = lab == o
        if not bool(m.any()):
            vals.append(torch.full((qry.shape[0],), 1e20, dtype=torch.float64, device=qry.device))
            idxs.append(torch.full((qry.shape[0],), -1, dtype=torch.int64, device=qry.device))
            continue
        dd = d[:, m]
        v, i = dd.min(dim=1)
        vals.append(v)
        idxs.append(torch.nonzero(m)[:, 0][i])
    return torch.stack(vals, 1), torch.stack(idxs, 1)


def run(api, ref, qry, lab, n_obj, engine="fr"):
    api.FORCE_SIMT_ENGINE, api.FORCE_EXACT3_ENGINE, api.FORCE_FR_ENGINE = engine == "simt", engine == "exact3", engine == "fr"
    try:
        out, _ = api.nearest_neighbor_features_per_object(ref.unsqueeze(1), qry.unsqueeze(1), lab.view(-1, 1, 1), 1,
                                                          torch.tensor(n_obj - 1))
    finally:
        api.FORCE_SIMT_ENGINE = api.FORCE_EXACT3_ENGINE = api.FORCE_FR_ENGINE = False
    return out.reshape(qry.shape[0], n_obj)


def tied_case(seed, R=6000, M=1500, C=100, n_obj=4, distinct=37, jitter=0.0):
    """References drawn from `distinct` prototype vectors: hundreds of exact (or, with jitter, near) ties per object."""
    gen = torch.Generator().manual_seed(seed)
    protos = 0.1 * torch.relu(torch.randn(distinct, C, generator=gen))
    pick = torch.randint(0, distinct, (R,), generator=gen)
    ref = protos[pick] + jitter * torch.randn(R, C, generator=gen)
    qry = protos[torch.randint(0, distinct, (M,), generator=gen)] + 0.01 * torch.randn(M, C, generator=gen)
    lab = torch.randint(0, n_obj, (R,), generator=gen).int()
    return ref.cuda(), qry.cuda(), lab.cuda(), n_obj


@pytest.mark.parametrize("jitter", [0.0, 1e-6, 1e-4])
def test_tied_references_all_paths(api, jitter):
    ref, qry, lab, n_obj = tied_case(3, jitter=jitter)
    want, _ = brute_force(ref, qry, lab, n_obj)
    got = run(api, ref, qry, lab, n_obj)
    assert rel_err(got, want) <= RAW_RTOL
    assert rel_err(run(api, ref, qry, lab, n_obj, "simt"), want) <= RAW_RTOL
    assert rel_err(run(api, ref, qry, lab, n_obj, "exact3"), want) <= RAW_RTOL
    assert torch.equal(got, run(api, ref, qry, lab, n_obj)), "deterministic"


def test_rescan_list_overflow_falls_back_in_place(api, option):
    ref, qry, lab, n_obj = tied_case(5)
    want = run(api, ref, qry, lab, n_obj)
    option("gm_fr_rescan_cap", 7)            # nearly every segment half is tied: the list overflows at once
    got = run(api, ref, qry, lab, n_obj)
    assert torch.equal(got, want)


@pytest.mark.parametrize("seg_tiles", [2, 3, 5])
def test_multi_tile_segments_equal_single_tile_segments(api, option, seg_tiles):
    """Large reference sets use segments of several tiles (1080p memory frames); forced here at a small size.  A pair's exact
    distance has one value whichever path evaluates it, so the results are bit-identical."""
    gen = torch.Generator().manual_seed(11)
    C, R, M, n_obj = 100, 9000, 2000, 5
    ref = (0.1 * torch.relu(torch.randn(R, C, generator=gen))).cuda()
    qry = (0.1 * torch.relu(torch.randn(M, C, generator=gen))).cuda()
    lab = torch.randint(-1, n_obj, (R,), generator=gen).int().cuda()
    lab[lab == 3] = 0                                                        # object 3 absent
    want = run(api, ref, qry, lab, n_obj)
    bf, _ = brute_force(ref, qry, lab, n_obj)
    assert rel_err(want[:, [0, 1, 2, 4]], bf[:, [0, 1, 2, 4]]) <= RAW_RTOL
    assert bool((want[:, 3] == 1e20).all())
    option("gm_fr_seg_tiles", seg_tiles)
    assert torch.equal(run(api, ref, qry, lab, n_obj), want)
    # ties across the tiles of a segment
    ref_t, qry_t, lab_t, n_t = tied_case(9, R=9000, M=1200)
    option("gm_fr_seg_tiles", 0)
    want_t = run(api, ref_t, qry_t, lab_t, n_t)
    option("gm_fr_seg_tiles", seg_tiles)
    assert torch.equal(run(api, ref_t, qry_t, lab_t, n_t), want_t)


@pytest.mark.parametrize("scale", [1e-4, 1.0, 3e3])
@pytest.mark.parametrize("C", [100, 64, 37, 128])
def test_scales_and_channel_layouts(api, scale, C):
    """Bias folded into the GEMM (C % 16 in 1..4), added in the epilogue (otherwise), one and two K blocks; operand scales
    far from 1 (the power-of-two scaling and the error bound are relative)."""
    gen = torch.Generator().manual_seed(C)
    R, M, n_obj = 2500, 900, 3
    ref = (scale * torch.randn(R, C, generator=gen)).cuda()
    qry = (scale * torch.randn(M, C, generator=gen)).cuda()
    lab = torch.randint(0, n_obj, (R,), generator=gen).int().cuda()
    want, _ = brute_force(ref, qry, lab, n_obj)
    got = run(api, ref, qry, lab, n_obj)
    assert float(((got.double() - want).abs() / want.abs().clamp_min(scale * scale)).max()) <= RAW_RTOL


def test_argmin_entry_matches_brute_force(api):
    from cvpr2020_manet_b200 import _lib
    from cvpr2020_manet_b200._device import stream_ptr, workspace
    gen = torch.Generator().manual_seed(21)
    C, R, M, n_obj = 100, 5000, 1300, 4
    ref = (0.1 * torch.relu(torch.randn(R, C, generator=gen))).cuda()
    qry = (0.1 * torch.relu(torch.randn(M, C, generator=gen))).cuda()
    lab = torch.randint(0, n_obj, (R,), generator=gen).int().cuda()
    lab[lab == 2] = 1                                                        # object 2 absent
    ref[100] = ref[50]                                                       # an exact duplicate pair with equal labels:
    lab[100] = lab[50]                                                       # ties resolve to the lowest index
    L = _lib.lib()
    dev = qry.device
    out = torch.empty(M, n_obj, device=dev)
    idx = torch.empty(M, n_obj, dtype=torch.int32, device=dev)
    ws = workspace(dev, L.manet_global_match_workspace_bytes(M, R, C, n_obj, 1), "global")
    _lib.check(L.manet_global_match_argmin_ws(ref.data_ptr(), C, 1, R, lab.data_ptr(), qry.data_ptr(), C, 1, M, C, n_obj, 0,
                                              out.data_ptr(), idx.data_ptr(), ws.data_ptr(), ws.numel(), stream_ptr(dev)), "argmin_ws")
    want, widx = brute_force(ref, qry, lab, n_obj)
    present = [0, 1, 3]
    assert rel_err(out[:, present], want[:, present]) <= RAW_RTOL
    assert bool((out[:, 2] == 1e20).all()) and bool((idx[:, 2] == -1).all())
    # the chosen reference realises the minimum (float64 check), and never is index 100 (its twin 50 comes first)
    d_at = ((qry.double().unsqueeze(1) - ref.double()[idx[:, present].long()]) ** 2).sum(-1)
    assert float(((d_at - want[:, present]).abs() / want[:, present].clamp_min(1.0)).max()) <= RAW_RTOL
    assert not bool((idx == 100).any())
    same = idx[:, present].long() == widx[:, present]
    assert float(same.float().mean()) > 0.999                              # differences only where two distances agree to fp32 noise


def test_tiny_and_degenerate_inputs(api):
    C = 100
    gen = torch.Generator().manual_seed(2)
    # one reference pixel; all-equal embeddings; a single query
    ref = (0.1 * torch.randn(1, C, generator=gen)).cuda()
    qry = (0.1 * torch.randn(5, C, generator=gen)).cuda()
    lab = torch.zeros(1, dtype=torch.int32).cuda()
    want, _ = brute_force(ref, qry, lab, 2)
    got = run(api, ref, qry, lab, 2)
    assert rel_err(got[:, 0], want[:, 0]) <= RAW_RTOL and bool((got[:, 1] == 1e20).all())
    ref = torch.full((700, C), 0.25).cuda()
    qry = torch.full((300, C), 0.25).cuda()
    lab = torch.randint(0, 3, (700,), generator=gen).int().cuda()
    assert bool((run(api, ref, qry, lab, 3) == 0).all())                   # exact zeros (difference form), thousands of ties
    zeros = torch.zeros(400, C).cuda()
    assert bool((run(api, zeros, zeros[:50], lab[:400], 3) == 0).all())


def test_reference_cache_gives_identical_results(api):
    """MANET_GM_REUSE_REF: along a propagation only the query changes; the cached reference side must give bit-identical
    maps, also when the query's power-of-two scale changes between frames (the bias is rebuilt for it) and with the fused
    normalise + memory epilogue; a changed reference (in-place edit bumps the version counter) forces a rebuild."""
    gen = torch.Generator().manual_seed(31)
    C, H, W, n_obj = 100, 40, 54, 4
    ref = (0.1 * torch.relu(torch.randn(C, H, W, generator=gen))).cuda().permute(1, 2, 0)
    lab = torch.randint(-1, n_obj, (H, W, 1), generator=gen).int().cuda()
    from cvpr2020_manet_b200.config import cfg
    saved = cfg.TEST_MODE
    cfg.TEST_MODE = True
    try:
        for forced in ("fr", "exact3", "auto"):       # the cache serves both tensor engines (auto picks by reference size)
            api.FORCE_FR_ENGINE, api.FORCE_EXACT3_ENGINE = forced == "fr", forced == "exact3"
            _reference_cache_case(api, ref, lab, C, H, W, n_obj, gen)
    finally:
        api.FORCE_FR_ENGINE = api.FORCE_EXACT3_ENGINE = False
        cfg.TEST_MODE = saved


def _reference_cache_case(api, ref, lab, C, H, W, n_obj, gen):
    cache = api.ReferenceOperands()
    lab = lab.clone()
    for _ in range(1):
        for i, scale in enumerate((1.0, 1.0, 37.0, 0.01, 1.0)):
            qry = (scale * 0.1 * torch.relu(torch.randn(C, H, W, generator=gen))).cuda().permute(1, 2, 0)
            want, _ = api.nearest_neighbor_features_per_object(ref, qry, lab, 1, torch.tensor(n_obj - 1))
            got, _ = api.nearest_neighbor_features_per_object(ref, qry, lab, 1, torch.tensor(n_obj - 1), reference_cache=cache)
            assert torch.equal(got, want), i
            mem_a = torch.rand(H, W, n_obj, 1, generator=gen).cuda()
            mem_b = mem_a.clone()
            want, _ = api.nearest_neighbor_features_per_object(ref, qry, lab, 1, torch.tensor(n_obj - 1), normalize=True, memory_frame=mem_a)
            got, _ = api.nearest_neighbor_features_per_object(ref, qry, lab, 1, torch.tensor(n_obj - 1), normalize=True, memory_frame=mem_b,
                                                              reference_cache=cache)
            assert torch.equal(got, want) and torch.equal(mem_a, mem_b), i
        lab[:5] = 0                                  # in-place edit: the cache must notice and rebuild
        qry = (0.1 * torch.relu(torch.randn(C, H, W, generator=gen))).cuda().permute(1, 2, 0)
        want, _ = api.nearest_neighbor_features_per_object(ref, qry, lab, 1, torch.tensor(n_obj - 1))
        got, _ = api.nearest_neighbor_features_per_object(ref, qry, lab, 1, torch.tensor(n_obj - 1), reference_cache=cache)
        assert torch.equal(got, want)
