"""DynamicSegHead (reference networks/IntVOS.py:488-525 and the feature assembly :663-671).

CPU part: the oracle restatement against golden vectors recorded from the UNMODIFIED reference class
(tests/golden/make_golden_seghead.py) and the host-side contract of the drop-in module.
GPU part (-m gpu): the sm_100a kernels through the C ABI against those goldens, against the oracle at the
480p / 5-object size, and on edge shapes.

Tolerance: logits |a-b| <= 2e-5 * max(1, max|b|).  The depthwise convs are fp32; the 1x1 convs run as fp16 hi/lo
split tensor-core products (22+ significant bits per operand, fp32 accumulation) -- fp32 grade, like the matchers.
"""
import numpy as np
import pytest
import torch

from oracle import manet_oracle as O

LOGIT_RTOL = 2e-5


def load_state(g):
    return {k[2:]: torch.from_numpy(g[k]) for k in g if k.startswith("p:")}


def logit_err(got, want):
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    return float(np.abs(got - want).max() / max(1.0, np.abs(want).max()))


# ------------------------------------------------------------------------------------------------ CPU
def test_oracle_seghead_matches_reference(golden):
    g = golden("seghead_ref")
    x = O.seghead_features(torch.from_numpy(g["cur"]), torch.from_numpy(g["gmap"]), torch.from_numpy(g["lmap"]),
                           torch.from_numpy(g["prev"]), torch.from_numpy(g["ids"]))
    assert x.shape == (3, 103, 21, 37)
    y = O.dynamic_seghead_forward(load_state(g), x)
    assert np.array_equal(y.numpy(), g["y"])          # same torch ops on the same host: bit-exact


def test_oracle_prop_seghead_matches_reference(golden):
    """The reference's IntVOS.prop_seghead end to end (matching + memories + head) against the oracle pieces."""
    g, gh = golden("prop_seghead_ref"), golden("seghead_ref")
    state = load_state(gh)
    embs = torch.from_numpy(g["embs"])
    scr = torch.from_numpy(g["scribble"])
    nobj, d = int(g["n_obj"]), int(g["d"])
    gm, lm = {}, ({}, {})
    ids = torch.arange(nobj + 1, dtype=torch.int32)
    for f in (1, 2):
        pl = torch.from_numpy(g[f"f{f}_prev_mask"])
        gmap, lmap = O.prop_matching_step(embs[0], embs[f - 1], embs[f], scr, pl, nobj, 1, d, True, gm, lm, "s", f, 1, 0)
        pred = O.dynamic_seghead_forward(state, O.seghead_features(embs[f], gmap, lmap, pl, ids)).permute(1, 0, 2, 3)
        assert logit_err(pred.numpy(), g[f"f{f}_pred"]) <= 1e-6


def test_module_contract():
    from cvpr2020_manet_b200.networks.seghead import DynamicSegHead, param_names
    head = DynamicSegHead()
    assert param_names() == O.seghead_param_names()
    ref_shapes = {"layer1.conv1.weight": (103, 1, 7, 7), "layer1.conv2.weight": (256, 103, 1, 1),
                  "layer4.conv1.weight": (256, 1, 7, 7), "layer4.bn2.running_var": (256,), "conv.weight": (1, 256, 1, 1),
                  "conv.bias": (1,)}
    sd = head.state_dict()
    for k, shp in ref_shapes.items():
        assert tuple(sd[k].shape) == shp
    head.eval()
    with pytest.raises(TypeError):                    # CPU tensors: no fallback
        head(torch.zeros(1, 103, 8, 8))
    with pytest.raises(ValueError):
        DynamicSegHead(in_dim=103, embed_dim=128)


# ------------------------------------------------------------------------------------------------ GPU
@pytest.fixture(scope="module")
def gpu_head(golden):
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from cvpr2020_manet_b200.networks.seghead import DynamicSegHead
    g = golden("seghead_ref")
    head = DynamicSegHead()
    head.load_state_dict(load_state(g), strict=False)
    return head.cuda().eval(), g


@pytest.mark.gpu
def test_gpu_seghead_golden(gpu_head):
    head, g = gpu_head
    x = O.seghead_features(torch.from_numpy(g["cur"]), torch.from_numpy(g["gmap"]), torch.from_numpy(g["lmap"]),
                           torch.from_numpy(g["prev"]), torch.from_numpy(g["ids"]))
    y = head(x.cuda())
    assert y.shape == g["y"].shape
    assert logit_err(y.cpu().numpy(), g["y"]) <= LOGIT_RTOL
    # non-contiguous input (channels-last storage) gives the same result
    y2 = head(x.cuda().permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2))
    assert torch.equal(y, y2)
    # fed by its parts: no repeat/cat (IntVOS.py:663-670)
    y3 = head.forward_parts(torch.from_numpy(g["cur"]).cuda(), torch.from_numpy(g["gmap"]).cuda(),
                            torch.from_numpy(g["lmap"]).cuda(), torch.from_numpy(g["prev"]).cuda(),
                            torch.from_numpy(g["ids"]).cuda())
    assert torch.equal(y, y3)


@pytest.mark.gpu
def test_gpu_seghead_training_mode_raises(gpu_head):
    head, _ = gpu_head
    head.train()
    try:
        with pytest.raises(NotImplementedError):
            head(torch.zeros(1, 103, 8, 8, device="cuda"))
    finally:
        head.eval()


@pytest.mark.gpu
@pytest.mark.parametrize("n,h,w,scale", [(1, 5, 7, 1.0), (2, 8, 16, 1.0), (1, 9, 17, 1.0), (3, 30, 54, 1e3), (2, 16, 33, 1e-4),
                                         (1, 12, 20, 0.0)])
def test_gpu_seghead_edge_shapes(gpu_head, n, h, w, scale):
    """Tiles smaller than / ragged against the 8x16 pixel unit, one object, large and tiny dynamic range (the
    per-pixel operand scale), all-zero input."""
    head, g = gpu_head
    gen = torch.Generator().manual_seed(h * 100 + w)
    x = scale * torch.randn(n, 103, h, w, generator=gen)
    want = O.dynamic_seghead_forward(load_state(g), x)
    got = head(x.cuda())
    assert logit_err(got.cpu().numpy(), want.numpy()) <= LOGIT_RTOL


@pytest.mark.gpu
def test_gpu_seghead_repacks_after_weight_update(gpu_head):
    head, g = gpu_head
    x = torch.randn(1, 103, 10, 18, generator=torch.Generator().manual_seed(5))
    state = load_state(g)
    before = head(x.cuda()).clone()
    saved = head.conv.bias.detach().clone()
    try:
        with torch.no_grad():
            head.conv.bias.add_(1.5)
        after = head(x.cuda())
        assert torch.allclose(after, before + 1.5, atol=1e-4)
    finally:
        with torch.no_grad():
            head.conv.bias.copy_(saved)
    assert torch.equal(head(x.cuda()), before)
    assert logit_err(before.cpu().numpy(), O.dynamic_seghead_forward(state, x).numpy()) <= LOGIT_RTOL


@pytest.mark.gpu
def test_gpu_seghead_480p_five_objects(gpu_head):
    """BASELINE shape: [6, 103, 120, 214] against the oracle (torch CPU fp32)."""
    head, g = gpu_head
    gen = torch.Generator().manual_seed(3)
    cur = 0.1 * torch.relu(torch.randn(100, 120, 214, generator=gen))
    gmap = torch.rand(1, 120, 214, 6, 1, generator=gen)
    lmap = torch.rand(1, 120, 214, 6, 1, generator=gen)
    prev = torch.randint(0, 6, (15, 27), generator=gen).repeat_interleave(8, 0).repeat_interleave(8, 1)[:120, :214].int()
    ids = torch.arange(6, dtype=torch.int32)
    want = O.dynamic_seghead_forward(load_state(g), O.seghead_features(cur, gmap, lmap, prev, ids))
    got = head.forward_parts(cur.cuda(), gmap.cuda(), lmap.cuda(), prev.cuda(), ids.cuda())
    assert got.shape == (6, 1, 120, 214)
    assert logit_err(got.cpu().numpy(), want.numpy()) <= LOGIT_RTOL
    # determinism (the two column halves of the last layer meet in one atomicAdd pair per logit)
    again = head.forward_parts(cur.cuda(), gmap.cuda(), lmap.cuda(), prev.cuda(), ids.cuda())
    assert torch.equal(got, again)


@pytest.mark.gpu
def test_gpu_prop_seghead_matches_reference(golden, gpu_head):
    """engine.prop_seghead (the reference's signature, IntVOS.py:583-681) against the reference's own run."""
    from cvpr2020_manet_b200 import engine
    from cvpr2020_manet_b200.config import cfg
    head, _ = gpu_head
    g = golden("prop_seghead_ref")
    saved = (cfg.TEST_MODE, cfg.MODEL_MAX_LOCAL_DISTANCE)
    cfg.TEST_MODE, cfg.MODEL_MAX_LOCAL_DISTANCE = True, int(g["d"])
    try:
        embs = torch.from_numpy(g["embs"]).cuda()
        scr = torch.from_numpy(g["scribble"]).cuda()
        h, w = scr.shape
        gm, lm = {}, ({}, {})
        for f in (1, 2):
            pl = torch.from_numpy(g[f"f{f}_prev_mask"]).cuda()
            res, gm, lm = engine.prop_seghead(ref_frame_embedding=embs[0:1], previous_frame_embedding=embs[f - 1:f],
                                              current_frame_embedding=embs[f:f + 1],
                                              ref_scribble_label=scr.view(1, 1, h, w).float(),
                                              previous_frame_mask=pl.view(1, 1, h, w).float(), seq_names=["s"],
                                              gt_ids=torch.tensor([int(g["n_obj"])]), k_nearest_neighbors=1,
                                              global_map_tmp_dic=gm, local_map_dics=lm, interaction_num=1,
                                              start_annotated_frame=0, frame_num=[f], dynamic_seghead=head)
            assert res["s"].shape == g[f"f{f}_pred"].shape
            # the maps entering the head are within 1e-5 of the reference's; the head's Lipschitz factor is O(10)
            assert logit_err(res["s"].cpu().numpy(), g[f"f{f}_pred"]) <= 2e-4
    finally:
        cfg.TEST_MODE, cfg.MODEL_MAX_LOCAL_DISTANCE = saved


@pytest.mark.gpu
def test_gpu_propagation_sequence_with_head(golden, gpu_head):
    """BASELINE config 3 in miniature: two interaction rounds of forward propagation over a short sequence through
    engine.prop_seghead (matching + both memories + head), against the oracle driven in lock step.  The previous-frame
    labels of both sides are the ORACLE's argmax, so every frame compares like with like."""
    from cvpr2020_manet_b200 import engine
    from cvpr2020_manet_b200.config import cfg
    head, gh = gpu_head
    state = load_state(gh)
    saved = (cfg.TEST_MODE, cfg.MODEL_MAX_LOCAL_DISTANCE)
    c, h, w, nobj, d, T = 100, 20, 28, 2, 3, 6
    cfg.TEST_MODE, cfg.MODEL_MAX_LOCAL_DISTANCE = True, d
    try:
        gen = torch.Generator().manual_seed(99)
        base = 0.1 * torch.relu(torch.randn(c, h, w, generator=gen))
        embs = torch.stack([base + 0.02 * t * torch.randn(c, h, w, generator=gen) for t in range(T)])
        ids = torch.arange(nobj + 1, dtype=torch.int32)
        gm_o, lm_o, gm_g, lm_g = {}, ({}, {}), {}, ({}, {})
        embs_g = embs.cuda()
        for rnd, ann in ((1, 0), (2, 1)):
            scr = torch.full((h, w), -1, dtype=torch.int32)
            scr[2 + rnd, 3:15] = 0
            scr[10, 5 + rnd:20] = 1
            if rnd == 1:
                scr[14:18, 9] = 2                       # object 2 absent from the second round's scribble
            prev_lab = torch.randint(0, nobj + 1, (h // 2, w // 2), generator=gen).repeat_interleave(2, 0).repeat_interleave(2, 1).int()
            for f in range(ann + 1, T):
                gmap, lmap = O.prop_matching_step(embs[ann], embs[f - 1], embs[f], scr, prev_lab, nobj, 1, d, True, gm_o, lm_o,
                                                  "s", f, rnd, ann)
                want = O.dynamic_seghead_forward(state, O.seghead_features(embs[f], gmap, lmap, prev_lab, ids)).permute(1, 0, 2, 3)
                res, gm_g, lm_g = engine.prop_seghead(ref_frame_embedding=embs_g[ann:ann + 1],
                                                      previous_frame_embedding=embs_g[f - 1:f],
                                                      current_frame_embedding=embs_g[f:f + 1],
                                                      ref_scribble_label=scr.cuda().view(1, 1, h, w).float(),
                                                      previous_frame_mask=prev_lab.cuda().view(1, 1, h, w).float(),
                                                      seq_names=["s"], gt_ids=torch.tensor([nobj]), k_nearest_neighbors=1,
                                                      global_map_tmp_dic=gm_g, local_map_dics=lm_g, interaction_num=rnd,
                                                      start_annotated_frame=ann, frame_num=[f], dynamic_seghead=head)
                assert logit_err(res["s"].cpu().numpy(), want.numpy()) <= 2e-4, (rnd, f)
                prev_lab = want[0].argmax(0).int()
        assert float((gm_g["s"].cpu() - gm_o["s"]).abs().max()) <= 1e-5
        assert float((lm_g[0]["s"].cpu() - lm_o[0]["s"]).abs().max()) <= 1e-5
        assert torch.equal(lm_g[1]["s"].cpu(), lm_o[1]["s"])
    finally:
        cfg.TEST_MODE, cfg.MODEL_MAX_LOCAL_DISTANCE = saved


@pytest.mark.gpu
def test_gpu_seghead_gemm_variants(golden):
    """The opt-in 1x1-conv kernels -- CTA pairs (cta_group::2, MANET_SH_PW_PAIR=1) and weight stages multicast inside a
    cluster of 2 or 4 (MANET_SH_PW_CLUSTER) -- give the same logits as the default kernel.  The switches are read once
    per process, so each variant runs in a subprocess."""
    import os
    import subprocess
    import sys
    code = (
        "import numpy as np, torch, sys\n"
        "sys.path.insert(0, '.')\n"
        "from cvpr2020_manet_b200.networks.seghead import DynamicSegHead\n"
        "g = dict(np.load('tests/golden/seghead_ref.npz'))\n"
        "head = DynamicSegHead(); head.load_state_dict({k[2:]: torch.from_numpy(v) for k, v in g.items() if k.startswith('p:')}, strict=False)\n"
        "head = head.cuda().eval()\n"
        "x = torch.randn(3, 103, 37, 70, generator=torch.Generator().manual_seed(1)).cuda()\n"
        "np.save(sys.argv[1], head(x).cpu().numpy())\n")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    outs = {}
    for name, env_add in (("default", {}), ("pair", {"MANET_SH_PW_PAIR": "1"}), ("mcast2", {"MANET_SH_PW_CLUSTER": "2"}),
                          ("mcast4", {"MANET_SH_PW_CLUSTER": "4"})):
        path = os.path.join(root, "gpurun_out", f"_gemm_variant_{name}.npy")
        os.makedirs(os.path.dirname(path), exist_ok=True)
        env = {k: v for k, v in os.environ.items() if k not in ("MANET_SH_PW_PAIR", "MANET_SH_PW_CLUSTER")}
        env.update(env_add)
        subprocess.run([sys.executable, "-c", code, path], check=True, cwd=root, env=env, timeout=300)
        outs[name] = np.load(path)
    for name in ("pair", "mcast2", "mcast4"):
        assert logit_err(outs[name], outs["default"]) <= 2e-6, name


# ------------------------------------------------------------------------------------------------ frame glue (8f-3)
def test_oracle_nearest_index_formula():
    """The kernel's nearest-downscale index, min(int(floorf(dst * (float)in/out)), in-1), is ATen's (IntVOS.py:598)."""
    for n_in, n_out in ((480, 120), (854, 214), (37, 9), (100, 33)):
        lab = torch.arange(n_in, dtype=torch.float32).view(1, 1, 1, n_in)
        want = torch.nn.functional.interpolate(lab, size=(1, n_out), mode="nearest").view(-1).long().numpy()
        scale = np.float32(n_in) / np.float32(n_out)
        got = np.minimum(np.floor(np.arange(n_out, dtype=np.float32) * scale).astype(np.int64), n_in - 1)
        assert np.array_equal(got, want)


@pytest.mark.gpu
@pytest.mark.parametrize("n,h,w,hf,wf", [(6, 120, 214, 480, 854), (3, 21, 37, 83, 149), (1, 5, 7, 5, 7), (4, 16, 20, 61, 33)])
def test_gpu_upsample_argmax(n, h, w, hf, wf):
    """Labels bit-exact against torch (interpolate bilinear align_corners + argmax, then nearest) wherever the two best
    upsampled logits differ by more than rounding noise; the small map equals the nearest-downscaled full map exactly."""
    from cvpr2020_manet_b200 import engine
    gen = torch.Generator().manual_seed(n * 1000 + h)
    pred = torch.randn(1, n, h, w, generator=gen) + 2.0 * torch.randn(1, n, h // 4 + 1, w // 4 + 1, generator=gen).repeat_interleave(4, 2).repeat_interleave(4, 3)[:, :, :h, :w]
    want, want_small, up = O.upsample_argmax(pred, (hf, wf))
    full, small = engine.upsample_argmax(pred.cuda(), (hf, wf))
    assert full.dtype == torch.int64 and full.shape == (1, hf, wf) and small.dtype == torch.int32
    top2 = torch.topk(up, min(2, n), dim=1).values
    decided = (top2[:, 0] - top2[:, -1] > 1e-5) if n > 1 else torch.ones_like(want, dtype=torch.bool)
    assert torch.equal(full.cpu()[decided], want[decided])
    assert float((full.cpu() != want).float().mean()) < 1e-3
    down = torch.nn.functional.interpolate(full.cpu().unsqueeze(0).float(), size=(h, w), mode="nearest").int()[0, 0]
    assert torch.equal(small.cpu(), down)


@pytest.mark.gpu
def test_gpu_propagate_sequence(golden, gpu_head):
    """engine.propagate_sequence (the loop of test.py:237-259 on the device) against the oracle driven frame by frame with
    the same label feedback; frames whose label maps agree keep the two sides in lock step."""
    from cvpr2020_manet_b200 import engine
    from cvpr2020_manet_b200.config import cfg
    from cvpr2020_manet_b200.networks.seghead import DynamicSegHead
    _, gh = gpu_head
    state = load_state(gh)
    # a freshly initialised head is a nearly constant function (kaiming fan_out depthwise taps are ~0.016: the per-object
    # signal decays tenfold per block), so its argmax is rounding noise.  Scale the taps and the three per-object input
    # channels up so that the argmax is decided over > 99 % of the frame (checked against the oracle below).
    state = {k: v.clone() for k, v in state.items()}
    for layer in range(1, 5):
        state[f"layer{layer}.conv1.weight"] *= 10.0
    state["layer1.conv2.weight"][:, 100:103] *= 20.0
    head = DynamicSegHead()
    head.load_state_dict(state, strict=False)
    head = head.cuda().eval()
    saved = (cfg.TEST_MODE, cfg.MODEL_MAX_LOCAL_DISTANCE)
    c, h, w, nobj, d, T, size = 100, 20, 28, 2, 3, 5, (79, 111)
    cfg.TEST_MODE, cfg.MODEL_MAX_LOCAL_DISTANCE = True, d
    try:
        gen = torch.Generator().manual_seed(5)
        base = 0.1 * torch.relu(torch.randn(c, h, w, generator=gen))
        embs = torch.stack([base + 0.02 * t * torch.randn(c, h, w, generator=gen) for t in range(T)])
        scr = torch.full((h, w), -1, dtype=torch.int32)
        scr[3, 3:15] = 0
        scr[10, 6:20] = 1
        scr[14:18, 9] = 2
        first = torch.randint(0, nobj + 1, (h // 2, w // 2), generator=gen).repeat_interleave(2, 0).repeat_interleave(2, 1).int()
        ids = torch.arange(nobj + 1, dtype=torch.int32)
        gm_g, lm_g = {}, ({}, {})
        got, last_small = engine.propagate_sequence(embs.cuda(), range(1, T), 0, scr.cuda(), first.cuda(), nobj, head, size,
                                                    gm_g, lm_g, "s", 1, d)
        gm_o, lm_o = {}, ({}, {})
        prev_small = first
        for f in range(1, T):
            gmap, lmap = O.prop_matching_step(embs[0], embs[f - 1], embs[f], scr, prev_small, nobj, 1, d, True, gm_o, lm_o, "s", f, 1, 0)
            pred = O.dynamic_seghead_forward(state, O.seghead_features(embs[f], gmap, lmap, prev_small, ids)).permute(1, 0, 2, 3)
            want, want_small, up = O.upsample_argmax(pred, size)
            # a random-init head separates the objects only weakly: compare where the two best upsampled logits differ by
            # more than the head's parity bound (2e-4 of the logit range), and require that to be most of the frame
            top2 = torch.topk(up, 2, dim=1).values
            decided = (top2[:, 0] - top2[:, 1]) > 4e-4 * float(up.abs().max())
            assert torch.equal(got[f].cpu()[decided], want[decided]), f
            assert float(decided.float().mean()) > 0.9, (f, float(decided.float().mean()))
            prev_small = got[f].cpu()                      # keep both sides on the device's labels
            prev_small = torch.nn.functional.interpolate(prev_small.unsqueeze(0).float(), size=(h, w), mode="nearest").int()[0, 0]
        assert torch.equal(last_small.cpu(), prev_small)
    finally:
        cfg.TEST_MODE, cfg.MODEL_MAX_LOCAL_DISTANCE = saved


def test_oracle_rough_roi_matches_reference():
    """oracle.rough_roi against the reference's own function (test.py:323-343), parsed out of the script when the
    reference tree is mounted (test.py as a whole imports packages this image lacks)."""
    import ast
    import os
    path = "/root/reference/test.py"
    if not os.path.exists(path):
        pytest.skip("reference tree not mounted")
    tree = ast.parse(open(path).read())
    fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "rough_ROI")
    ns = {"torch": torch}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), path, "exec"), ns)
    gen = torch.Generator().manual_seed(2)
    for h, w in ((120, 214), (30, 41), (64, 64)):
        lab = torch.full((2, 1, h, w), -1.0)
        for b in range(2):
            y, x = int(torch.randint(0, h, (1,), generator=gen)), int(torch.randint(0, w - 5, (1,), generator=gen))
            lab[b, 0, y, x:x + 5] = float(b + 1)
            lab[b, 0, min(h - 1, y + 7), min(w - 1, x + 2)] = 0.0
        assert torch.equal(O.rough_roi(lab), ns["rough_ROI"](lab))


@pytest.mark.gpu
@pytest.mark.parametrize("h,w,dtype", [(120, 214, torch.float32), (30, 41, torch.int32), (25, 300, torch.float32)])
def test_gpu_rough_roi(h, w, dtype):
    from cvpr2020_manet_b200 import engine
    gen = torch.Generator().manual_seed(h)
    lab = torch.full((3, 1, h, w), -1, dtype=dtype)
    for b in range(3):
        for o in range(3):
            y, x = int(torch.randint(0, h, (1,), generator=gen)), int(torch.randint(0, w - 4, (1,), generator=gen))
            lab[b, 0, y, x:x + 4] = o
    lab[2, 0, 0, 0] = 1
    lab[2, 0, h - 1, w - 1] = 2                         # box reaches every border
    got = engine.rough_ROI(lab.cuda())
    assert got.dtype == dtype
    assert torch.equal(got.cpu(), O.rough_roi(lab))


@pytest.mark.gpu
@pytest.mark.parametrize("in_dim", [8, 35, 128])
def test_gpu_seghead_other_input_widths(in_dim):
    """in_dim other than 103: whole zero chunks of the padded first layer (in_dim = 8, 35), no padding at all (128)."""
    from cvpr2020_manet_b200.networks.seghead import DynamicSegHead
    torch.manual_seed(in_dim)
    head = DynamicSegHead(in_dim=in_dim).eval()
    with torch.no_grad():
        for layer in (head.layer1, head.layer2, head.layer3, head.layer4):
            layer.conv1.weight.mul_(6.0)                 # keep some signal through the four blocks
            for bn in (layer.bn1, layer.bn2):
                bn.running_mean.normal_(0.0, 0.1)
                bn.running_var.uniform_(0.5, 1.5)
                bn.weight.uniform_(0.5, 1.5)
                bn.bias.normal_(0.0, 0.2)
    state = {k: v.detach().clone() for k, v in head.state_dict().items()}
    x = torch.randn(2, in_dim, 19, 45, generator=torch.Generator().manual_seed(1))
    want = O.dynamic_seghead_forward(state, x)
    got = head.cuda()(x.cuda())
    assert logit_err(got.cpu().numpy(), want.numpy()) <= LOGIT_RTOL


@pytest.mark.gpu
def test_gpu_int_seghead_matches_reference(golden):
    """engine.int_seghead (the reference's signature, IntVOS.py:683-764) against the reference's own run: the tensor handed
    to the interaction head is exact (embedding, scribble mask, previous-round mask), the memories within 1e-5."""
    from cvpr2020_manet_b200 import engine
    from cvpr2020_manet_b200.config import cfg
    g = golden("int_seghead_ref")
    saved = (cfg.TEST_MODE, cfg.MODEL_MAX_LOCAL_DISTANCE)
    cfg.TEST_MODE, cfg.MODEL_MAX_LOCAL_DISTANCE = True, int(g["d"])
    try:
        embs = torch.from_numpy(g["embs"]).cuda()
        _, c, h, w = embs.shape
        nobj = int(g["n_obj"])
        seen = []

        def head(x):
            seen.append(x)
            return torch.zeros(x.shape[0], 1, x.shape[2], x.shape[3], device=x.device)

        gm, lm = {}, ({}, {})
        for rnd, frame, first in ((1, 1, True), (2, 2, False)):
            scr = torch.from_numpy(g[f"r{rnd}_scribble"]).cuda()
            prev_round = torch.from_numpy(g[f"r{rnd}_prev_round"]).cuda()
            res, lm = engine.int_seghead(ref_frame_embedding=embs[frame:frame + 1], ref_scribble_label=scr.view(1, 1, h, w).float(),
                                         prev_round_label=None if first else prev_round.view(1, 1, h, w).float(),
                                         global_map_tmp_dic=gm, local_map_dics=lm, interaction_num=rnd, seq_names=["s"],
                                         gt_ids=torch.tensor([nobj]), frame_num=[frame], first_inter=first, inter_seghead=head)
            assert tuple(res["s"].shape) == tuple(g[f"r{rnd}_pred_shape"])
            assert np.array_equal(seen[-1].cpu().numpy(), g[f"r{rnd}_to_cat"])
        assert float(np.abs(gm["s"][:3].cpu().numpy() - g["final_global_mem"]).max()) <= 1e-5
        assert np.array_equal(lm[1]["s"][:3, :3].cpu().numpy(), g["final_local_dist"])
        assert bool((lm[0]["s"] == 1).all()) == bool(g["final_local_mem_is_ones"])
    finally:
        cfg.TEST_MODE, cfg.MODEL_MAX_LOCAL_DISTANCE = saved


def _int_to_cat(emb, scr, prev_round, ids):
    """to_cat of IntVOS.py:741-757 (embedding repeated per object, scribble mask, previous-round mask / first-round one-hot)."""
    rep = emb.unsqueeze(0).repeat(ids.numel(), 1, 1, 1)
    scr_mask = (scr.float().unsqueeze(-1) == ids.float()).permute(2, 0, 1).unsqueeze(1).float()
    if prev_round is None:
        prev_mask = torch.zeros_like(scr_mask)
        prev_mask[0] = 1.0
    else:
        prev_mask = (prev_round.float().unsqueeze(-1) == ids.float()).permute(2, 0, 1).unsqueeze(1).float()
    return torch.cat((rep, scr_mask, prev_mask), 1)


def test_oracle_default_interaction_head_matches_reference(golden):
    """The reference's default interaction head (config.py:52 -> IntVOS.py:554: DynamicSegHead(in_dim=C+2)) through its own
    int_seghead: the oracle's head on the restated to_cat reproduces the recorded logits."""
    g = golden("int_seghead_default_head_ref")
    state = load_state(g)
    embs = torch.from_numpy(g["embs"])
    ids = torch.arange(int(g["n_obj"]) + 1, dtype=torch.int32)
    for rnd, frame, first in ((1, 1, True), (2, 2, False)):
        scr = torch.from_numpy(g[f"r{rnd}_scribble"])
        prev = None if first else torch.from_numpy(g[f"r{rnd}_prev_round"])
        pred = O.dynamic_seghead_forward(state, _int_to_cat(embs[frame], scr, prev, ids)).permute(1, 0, 2, 3)
        assert logit_err(pred.numpy(), g[f"r{rnd}_pred"]) <= 1e-6


@pytest.mark.gpu
def test_gpu_int_seghead_default_head_matches_reference(golden):
    """engine.int_seghead with this package's DynamicSegHead(in_dim=C+2) as the interaction head -- the reference's default
    configuration, whole branch on the sm_100a kernels, head fed by its parts -- against the reference's own logits."""
    from cvpr2020_manet_b200 import engine
    from cvpr2020_manet_b200.config import cfg
    from cvpr2020_manet_b200.networks.seghead import DynamicSegHead
    g = golden("int_seghead_default_head_ref")
    saved = (cfg.TEST_MODE, cfg.MODEL_MAX_LOCAL_DISTANCE)
    cfg.TEST_MODE, cfg.MODEL_MAX_LOCAL_DISTANCE = True, int(g["d"])
    try:
        embs = torch.from_numpy(g["embs"]).cuda()
        _, c, h, w = embs.shape
        nobj = int(g["n_obj"])
        head = DynamicSegHead(in_dim=c + 2)
        head.load_state_dict(load_state(g), strict=False)
        head = head.cuda().eval()
        gm, lm = {}, ({}, {})
        for rnd, frame, first in ((1, 1, True), (2, 2, False)):
            scr = torch.from_numpy(g[f"r{rnd}_scribble"]).cuda()
            prev_round = torch.from_numpy(g[f"r{rnd}_prev_round"]).cuda()
            res, lm = engine.int_seghead(ref_frame_embedding=embs[frame:frame + 1], ref_scribble_label=scr.view(1, 1, h, w).float(),
                                         prev_round_label=None if first else prev_round.view(1, 1, h, w).float(),
                                         global_map_tmp_dic=gm, local_map_dics=lm, interaction_num=rnd, seq_names=["s"],
                                         gt_ids=torch.tensor([nobj]), frame_num=[frame], first_inter=first, inter_seghead=head)
            assert tuple(res["s"].shape) == g[f"r{rnd}_pred"].shape
            assert logit_err(res["s"].cpu().numpy(), g[f"r{rnd}_pred"]) <= LOGIT_RTOL, rnd
            # the parts entry equals the dense entry on the assembled tensor (same kernels after layer 1's loader)
            ids = torch.arange(nobj + 1, dtype=torch.int32).cuda()
            dense = head(_int_to_cat(embs[frame], scr, None if first else prev_round, ids))
            assert logit_err(dense.permute(1, 0, 2, 3).cpu().numpy(), g[f"r{rnd}_pred"]) <= LOGIT_RTOL
        assert float(np.abs(gm["s"][:3].cpu().numpy() - g["final_global_mem"]).max()) <= 1e-5
    finally:
        cfg.TEST_MODE, cfg.MODEL_MAX_LOCAL_DISTANCE = saved


@pytest.mark.gpu
def test_gpu_int_seghead_default_head_480p(golden):
    """The interaction head at the 480p / 5-object size against the CPU oracle (weights of the golden), both round forms."""
    from cvpr2020_manet_b200.networks.seghead import DynamicSegHead
    g = golden("int_seghead_default_head_ref")
    state = load_state(g)
    head = DynamicSegHead(in_dim=102)
    head.load_state_dict(state, strict=False)
    head = head.cuda().eval()
    gen = torch.Generator().manual_seed(17)
    emb = 0.1 * torch.relu(torch.randn(100, 120, 214, generator=gen))
    ids = torch.arange(6, dtype=torch.int32)
    scr = torch.full((120, 214), -1, dtype=torch.int32)
    for o in range(6):
        scr[10 + 17 * o, 20:160] = o
    prev = torch.randint(0, 6, (15, 27), generator=gen).repeat_interleave(8, 0).repeat_interleave(8, 1)[:120, :214].int()
    for prev_round in (None, prev):
        want = O.dynamic_seghead_forward(state, _int_to_cat(emb, scr, prev_round, ids))
        got = head.forward_parts_interaction(emb.cuda(), scr.cuda(), None if prev_round is None else prev_round.cuda(), ids.cuda())
        assert got.shape == (6, 1, 120, 214)
        assert logit_err(got.cpu().numpy(), want.numpy()) <= LOGIT_RTOL


@pytest.mark.gpu
def test_gpu_prop_seghead_under_grad_matches_reference(golden):
    """Training form (train_stage1.py:126): engine.prop_seghead with gradients flowing into the embeddings through both
    matchers (autograd functions over the sm_100a kernels) and a differentiable torch head, against pred and gradients
    recorded from the reference's own autograd.  Tolerance on gradients: 2e-4 of the largest one (as tests/test_gpu_autograd)."""
    from cvpr2020_manet_b200 import engine
    from cvpr2020_manet_b200.config import cfg
    g, gh = golden("prop_seghead_grad_ref"), golden("seghead_ref")
    state = {k: v.cuda() for k, v in load_state(gh).items()}
    saved = (cfg.TEST_MODE, cfg.MODEL_MAX_LOCAL_DISTANCE)
    cfg.TEST_MODE, cfg.MODEL_MAX_LOCAL_DISTANCE = False, int(g["d"])
    tf32 = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False     # the torch head must be fp32 like the reference's
    try:
        embs = torch.from_numpy(g["embs"]).cuda().requires_grad_(True)
        _, c, h, w = embs.shape
        ref_lab = torch.from_numpy(g["ref_label"]).cuda()
        prev_lab = torch.from_numpy(g["prev_label"]).cuda()
        weight = torch.from_numpy(g["weight"]).cuda()
        res = engine.prop_seghead(ref_frame_embedding=embs[0:1], previous_frame_embedding=embs[1:2],
                                  current_frame_embedding=embs[2:3], ref_scribble_label=ref_lab.view(1, 1, h, w).float(),
                                  previous_frame_mask=prev_lab.view(1, 1, h, w).float(), seq_names=["s"],
                                  gt_ids=torch.tensor([int(g["n_obj"])]), k_nearest_neighbors=1, global_map_tmp_dic=None,
                                  local_map_dics=None, dynamic_seghead=lambda x: O.dynamic_seghead_forward(state, x))
        pred = res["s"]
        assert logit_err(pred.detach().cpu().numpy(), g["pred"]) <= 2e-4
        (pred * weight).sum().backward()
        want = g["grad_embs"]
        err = float(np.abs(embs.grad.cpu().numpy() - want).max() / np.abs(want).max())
        assert err <= 2e-4, err
        # this package's inference-form head cannot sit on the training path: loud failure, not a silent gradient cut
        from cvpr2020_manet_b200.networks.seghead import DynamicSegHead
        with pytest.raises(NotImplementedError):
            engine.prop_seghead(ref_frame_embedding=embs[0:1], previous_frame_embedding=embs[1:2],
                                current_frame_embedding=embs[2:3], ref_scribble_label=ref_lab.view(1, 1, h, w).float(),
                                previous_frame_mask=prev_lab.view(1, 1, h, w).float(), seq_names=["s"],
                                gt_ids=torch.tensor([int(g["n_obj"])]), k_nearest_neighbors=1,
                                dynamic_seghead=DynamicSegHead().cuda().eval())
    finally:
        cfg.TEST_MODE, cfg.MODEL_MAX_LOCAL_DISTANCE = saved
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32


def test_reference_checkpoint_keys_load():
    """A state_dict of the reference's own DynamicSegHead (IntVOS.py:510-525) loads into the drop-in module: same parameter and
    buffer names and shapes (only BatchNorm's num_batches_tracked bookkeeping may differ)."""
    from oracle import ref_shim
    if not ref_shim.reference_available():
        pytest.skip("reference tree not mounted")
    from cvpr2020_manet_b200.networks.seghead import DynamicSegHead
    ref = ref_shim.load_reference()
    torch.manual_seed(3)
    theirs = ref.DynamicSegHead().state_dict()
    ours = DynamicSegHead()
    result = ours.load_state_dict(theirs, strict=False)
    assert [k for k in result.missing_keys if not k.endswith("num_batches_tracked")] == []
    assert [k for k in result.unexpected_keys if not k.endswith("num_batches_tracked")] == []
    mine = ours.state_dict()
    for k, v in theirs.items():
        if k in mine:
            assert torch.equal(mine[k], v), k
