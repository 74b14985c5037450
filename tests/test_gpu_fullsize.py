"""Full-size parity: the exact tensors bench.py times (480p embedding 100x120x214, N=6, d=12) and BASELINE
config 1 (N=3, d=9, the reference's own torch.rand distribution) against the CPU oracle -- both matching maps and
both map memories after two interaction rounds, through the device API and through the host-buffer C-ABI session
(the call bench.py's `e2e` measures).  The oracle needs ~2 s per step on the box's host cores.

Tolerances as everywhere: normalised maps <= 1e-5 absolute, sentinels (exactly 1.0 where the
oracle is exactly 1.0), local-map round selection and distance table bit-exact."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

MAP_ATOL = 1e-5


@pytest.fixture()
def cfg_guard():
    from cvpr2020_manet_b200.config import cfg
    saved = dict(vars(cfg))
    yield cfg
    for k, v in saved.items():
        setattr(cfg, k, v)


def _blob(gen, n_ids, H, W):
    return torch.randint(0, n_ids, (H // 8, W // 8 + 1), generator=gen).repeat_interleave(8, 0).repeat_interleave(8, 1)[:H, :W].int().contiguous()


def _two_round_inputs(kind):
    """(embeddings ref/prev/cur [C,H,W], per-round (ref_labels, prev_labels), n_ids, d)."""
    if kind == "headline":
        import bench
        ref, prev, cur, ref_lab, prev_lab = bench.synth_inputs(1000)          # rank 0's tensors in bench.py
        gen = torch.Generator().manual_seed(4242)
        n_ids, d = bench.N_IDS, bench.D_LOCAL
        H, W = bench.H, bench.W
        # round 2: a sparser scribble-like reference (object 4 absent -> sentinel) and new previous-frame labels
        ref2 = _blob(gen, n_ids, H, W)
        ref2[ref2 == 4] = 0
        rounds = [(ref_lab, prev_lab), (ref2, _blob(gen, n_ids, H, W))]
        return (ref, prev, cur), rounds, n_ids, d
    gen = torch.Generator().manual_seed(0)                                    # BASELINE config 1
    C, H, W, n_ids, d = 100, 120, 214, 3, 9
    ref, prev, cur = (torch.rand(C, H, W, generator=gen) for _ in range(3))
    rounds = [(_blob(gen, n_ids, H, W), _blob(gen, n_ids, H, W)), (_blob(gen, n_ids, H, W), _blob(gen, n_ids, H, W))]
    return (ref, prev, cur), rounds, n_ids, d


def _oracle_session(embs, rounds, n_ids, d, frame, starts):
    from oracle import manet_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    gmem, lmem = {}, ({}, {})
    outs = []
    for rnd, ((ref_lab, prev_lab), start) in enumerate(zip(rounds, starts), 1):
        g, l = O.prop_matching_step(embs[0], embs[1], embs[2], ref_lab, prev_lab, n_ids - 1, 1, d, True, gmem, lmem, "s", frame,
                                    rnd, start)
        outs.append((g.clone(), l.clone()))
    return outs, gmem["s"][frame], lmem[0]["s"][frame, :2], lmem[1]["s"][frame, :2]


def _check_map(got, want, what):
    got, want = got.reshape(-1).cpu().numpy(), want.reshape(-1).numpy()
    assert np.max(np.abs(got - want)) <= MAP_ATOL, what
    ones = want == 1.0
    assert np.array_equal(got[ones], want[ones]), what + ": sentinels must be exactly 1.0"


@pytest.mark.parametrize("kind", ["headline", "config1"])
def test_full_size_two_rounds_vs_oracle(kind, cfg_guard):
    from cvpr2020_manet_b200 import engine
    cfg_guard.TEST_MODE = True
    embs, rounds, n_ids, d = _two_round_inputs(kind)
    frame, starts = 7, (0, 3)           # round 2 is annotated nearer to the frame: its local map wins (IntVOS.py:654-659)
    want, want_gmem, want_lmem, want_ldist = _oracle_session(embs, rounds, n_ids, d, frame, starts)
    dev = [e.cuda() for e in embs]
    gmem, lmem = {}, ({}, {})
    for rnd, ((ref_lab, prev_lab), start) in enumerate(zip(rounds, starts), 1):
        g, l = engine.prop_matching_step(dev[0], dev[1], dev[2], ref_lab.cuda(), prev_lab.cuda(), n_ids - 1, 1, d, gmem, lmem,
                                         "s", frame, rnd, start)
        _check_map(g, want[rnd - 1][0], f"{kind} round {rnd} global map")
        _check_map(l, want[rnd - 1][1], f"{kind} round {rnd} local map")
    _check_map(gmem["s"][frame], want_gmem, f"{kind} global-map memory")
    _check_map(lmem[0]["s"][frame, :2], want_lmem, f"{kind} local-map memory")
    assert np.array_equal(lmem[1]["s"][frame, :2].cpu().numpy(), want_ldist.numpy())
    if kind == "headline":
        # object 4 is absent from round 2's reference: without the memory the raw distance is the sentinel, bit-exact
        from cvpr2020_manet_b200.networks import IntVOS
        raw, _ = IntVOS.nearest_neighbor_features_per_object(dev[0].permute(1, 2, 0), dev[2].permute(1, 2, 0),
                                                             rounds[1][0].cuda().unsqueeze(-1), 1, n_ids - 1)
        assert bool((raw[0, :, :, 4, 0] == 1e20).all()) and bool((raw[0, :, :, :4, 0] < 1e19).all())


def test_full_size_host_buffer_session_vs_oracle(cfg_guard):
    """The C-ABI session bench.py's `e2e` goes through (pinned host buffers in, both maps out), at the headline size."""
    from cvpr2020_manet_b200 import engine
    cfg_guard.TEST_MODE = True
    embs, rounds, n_ids, d = _two_round_inputs("headline")
    frame, starts = 7, (0, 3)
    want, want_gmem, _, _ = _oracle_session(embs, rounds, n_ids, d, frame, starts)
    C, H, W = embs[0].shape
    sess = engine.MatchingSession(H, W, C, n_ids, d, n_frames=16)
    try:
        sess.ref[:], sess.prev[:], sess.cur[:] = embs[0].numpy(), embs[1].numpy(), embs[2].numpy()
        for rnd, ((ref_lab, prev_lab), start) in enumerate(zip(rounds, starts), 1):
            sess.ref_labels[:], sess.prev_labels[:] = ref_lab.numpy(), prev_lab.numpy()
            og, ol = sess.step_host(frame, rnd, start)
            _check_map(torch.from_numpy(og.copy()), want[rnd - 1][0], f"session round {rnd} global map")
            _check_map(torch.from_numpy(ol.copy()), want[rnd - 1][1], f"session round {rnd} local map")
    finally:
        sess.close()
