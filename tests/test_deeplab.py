"""The embedding extractor / IntVOS wrapper (reference networks/deeplab.py:27-52, backbone/resnet.py, aspp.py,
decoder.py, IntVOS.py:527-581).  It is the caller of the hot path, written from the architecture in plain torch; what is
checked here (CPU): the topology equals the reference's -- identical state_dict keys and shapes, and identical outputs once
the reference's randomly initialised weights are loaded -- and the output geometry at the 480p shape."""
import pytest
import torch

from oracle import ref_shim


def _ours():
    from cvpr2020_manet_b200.networks import deeplab
    return deeplab


def _strip(keys):
    return sorted(k for k in keys if not k.endswith("num_batches_tracked"))


def test_output_geometry_small():
    net = _ours().DeepLab().eval()
    with torch.no_grad():
        y = net(torch.randn(1, 3, 64, 96))
    assert y.shape == (1, 256, 16, 24)
    n_params = sum(p.numel() for p in net.parameters())
    assert 58_000_000 < n_params < 61_000_000          # DeepLabv3+ / ResNet-101 without its classifier


@pytest.mark.skipif(not ref_shim.reference_available(), reason="reference tree not mounted")
def test_deeplab_matches_reference_topology_and_output():
    ref_shim.load_reference()                          # puts the stub `config` in place and /root/reference on sys.path
    from networks.deeplab import DeepLab as RefDeepLab  # the unmodified reference class
    torch.manual_seed(0)
    theirs = RefDeepLab(backbone="resnet", sync_bn=False).eval()
    theirs.decoder.cls_conv = torch.nn.Sequential()
    ours = _ours().DeepLab().eval()
    sd = theirs.state_dict()
    mine = ours.state_dict()
    assert _strip(sd) == _strip(mine)
    for k in _strip(sd):
        assert sd[k].shape == mine[k].shape, k
    ours.load_state_dict(sd, strict=False)
    x = torch.randn(2, 3, 65, 97, generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        a, b = theirs(x), ours(x)
    assert a.shape == b.shape == (2, 256, 17, 25)
    assert torch.allclose(a, b, rtol=1e-5, atol=1e-5 * float(a.abs().max()))


@pytest.mark.skipif(not ref_shim.reference_available(), reason="reference tree not mounted")
def test_intvos_wrapper_state_dict_matches_reference():
    mod = ref_shim.load_reference()
    from networks.deeplab import DeepLab as RefDeepLab
    torch.manual_seed(0)
    theirs = mod.IntVOS(mod.cfg, RefDeepLab(backbone="resnet", sync_bn=False))
    ours = _ours().IntVOS()
    sd, mine = theirs.state_dict(), ours.state_dict()
    assert _strip(sd) == _strip(mine)
    for k in _strip(sd):
        assert sd[k].shape == mine[k].shape, k
    result = ours.load_state_dict(sd, strict=False)
    assert [k for k in result.missing_keys if not k.endswith("num_batches_tracked")] == []
    # extract_feature: same embedding from the same weights (CPU torch on both sides)
    theirs.eval(); ours.eval()
    x = torch.randn(1, 3, 64, 96, generator=torch.Generator().manual_seed(2))
    with torch.no_grad():
        a, b = theirs.extract_feature(x), ours.extract_feature(x)
    assert a.shape == b.shape == (1, 100, 16, 24)
    assert torch.allclose(a, b, rtol=1e-5, atol=1e-5 * float(a.abs().max()))


@pytest.mark.gpu
def test_gpu_intvos_forward_matches_oracle_pipeline():
    """IntVOS.forward (IntVOS.py:556-575) on the GPU: embeddings from the torch extractor, then matching + DynamicSegHead on
    the sm_100a kernels, against the CPU oracle fed with the SAME embeddings (so the comparison isolates the hot path)."""
    import numpy as np
    from oracle import manet_oracle as O
    from cvpr2020_manet_b200.config import cfg
    from cvpr2020_manet_b200.networks import IntVOS as api
    deeplab = _ours()
    saved = (cfg.TEST_MODE, cfg.MODEL_MAX_LOCAL_DISTANCE)
    cfg.TEST_MODE, cfg.MODEL_MAX_LOCAL_DISTANCE = False, 4
    try:
        torch.manual_seed(0)
        model = deeplab.IntVOS(cfg).cuda().eval()
        gen = torch.Generator().manual_seed(3)
        x = torch.randn(3, 3, 96, 128, generator=gen).cuda()
        with torch.no_grad():
            # give the embedding a trained-network scale (unit-variance pre-activations), see bench.intvos_forward_leg
            pre = model.embedding_conv(model.relu1(model.bn1(model.seperate_conv(model.feature_extracter(x)))))
            model.bn2.running_mean.copy_(pre.mean(dim=(0, 2, 3)))
            model.bn2.running_var.copy_(pre.var(dim=(0, 2, 3)))
            emb = model.extract_feature(x)
        n_obj = 2
        lab = lambda: torch.randint(0, n_obj + 1, (12, 16), generator=gen).repeat_interleave(8, 0).repeat_interleave(8, 1)
        ref_lab, prev_lab = lab(), lab()
        with torch.no_grad():
            res = model(x, ref_lab.view(1, 1, 96, 128).float().cuda(), prev_lab.view(1, 1, 96, 128).float().cuda(),
                        seq_names=["s"], gt_ids=torch.tensor([n_obj]), k_nearest_neighbors=1)
        pred = res["s"].cpu()
        h, w = emb.shape[2:]
        assert pred.shape == (1, n_obj + 1, h, w) == (1, 3, 24, 32)
        e = emb.cpu()
        small = lambda t: torch.nn.functional.interpolate(t.view(1, 1, 96, 128).float(), size=(h, w), mode="nearest").int()[0, 0]
        g, l = O.prop_matching_step(e[0], e[1], e[2], small(ref_lab), small(prev_lab), n_obj, 1, 4, False)
        state = {k: v.detach().cpu() for k, v in model.dynamic_seghead.state_dict().items()}
        ids = torch.arange(n_obj + 1, dtype=torch.int32)
        want = O.dynamic_seghead_forward(state, O.seghead_features(e[2], g, l, small(prev_lab), ids)).permute(1, 0, 2, 3)
        err = float(np.abs(pred.numpy() - want.numpy()).max() / max(1.0, float(want.abs().max())))
        assert err <= 2e-4, err
        stats = api.local_match_guard_stats(h, w, 100, n_obj + 1, 4)
        assert stats["engine"].startswith(("tcgen05", "cuda-core")) and stats["threshold"] > 0
    finally:
        cfg.TEST_MODE, cfg.MODEL_MAX_LOCAL_DISTANCE = saved
