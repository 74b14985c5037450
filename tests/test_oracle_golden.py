"""The CPU oracle (oracle/manet_oracle.py, oracle/naive.py) against the golden
vectors produced by the unmodified reference (tests/golden/make_golden.py)."""
import numpy as np
import pytest
import torch

from oracle import manet_oracle as O
from oracle import naive

GLOBAL = ["global_k1_A", "global_k1_B_absent", "global_testmode_scribble", "global_k3_A",
          "global_multiframe_ref", "global_gtids_none", "global_k2_testmode"]
LOCAL = ["local_d3_even", "local_d4_odd", "local_d12_window_gt_image", "local_d9_A", "local_d2_scaled"]


def rel_close(a, b, tol=1e-5):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.max(np.abs(a - b) / np.maximum(1.0, np.abs(b))) <= tol


@pytest.mark.parametrize("name", GLOBAL)
def test_global_oracle_matches_reference(golden, name):
    g = golden(name)
    ref = torch.from_numpy(g["ref_chw"]).permute(1, 2, 0)
    qry = torch.from_numpy(g["query_chw"]).permute(1, 2, 0)
    lab = torch.from_numpy(g["labels"]).unsqueeze(-1)
    gt = torch.tensor(int(g["n_obj"])) if int(g["pass_gt"]) else None
    out, ids = O.global_match(ref, qry, lab, int(g["k"]), gt, n_chunks=int(g["n_chunks"]),
                              test_mode=bool(g["test_mode"]))
    assert out.shape == g["out"].shape
    assert ids.dtype == torch.int32 and np.array_equal(ids.numpy(), g["ids"])
    # same op sequence as the reference on the same host => bit-exact here
    assert np.array_equal(out.numpy(), g["out"])


@pytest.mark.parametrize("name", GLOBAL)
def test_global_naive_matches_reference(golden, name):
    g = golden(name)
    c = g["ref_chw"].shape[0]
    ref = g["ref_chw"].transpose(1, 2, 0).reshape(-1, c)
    qry = g["query_chw"].transpose(1, 2, 0).reshape(-1, c)
    n_ids = g["ids"].shape[0]
    out = naive.global_match(ref, qry, g["labels"].reshape(-1), n_ids, k=int(g["k"]),
                             drop_unlabelled=bool(g["test_mode"]))
    want = g["out"].reshape(-1, n_ids)
    assert rel_close(out, want, 2e-5)
    assert np.array_equal(want == np.float32(1e20), out == 1e20)


def test_selected_pixel_bit_exact(golden):
    g = golden("selected_pixel")
    l2, e2 = O.select_labelled(torch.from_numpy(g["labels"]), torch.from_numpy(g["emb"]))
    assert np.array_equal(l2.numpy(), g["out_labels"]) and np.array_equal(e2.numpy(), g["out_emb"])


@pytest.mark.parametrize("name", LOCAL)
def test_local_oracle_matches_reference(golden, name):
    g = golden(name)
    prev = torch.from_numpy(g["prev_chw"]).permute(1, 2, 0)
    cur = torch.from_numpy(g["cur_chw"]).permute(1, 2, 0)
    lab = torch.from_numpy(g["labels"]).unsqueeze(-1)
    out = O.local_match(prev, cur, lab, torch.from_numpy(g["ids"]), int(g["d"]))
    assert np.array_equal(out.numpy(), g["out"])
    if "window" in g:
        win = O.local_window_distances(cur, prev, int(g["d"]))
        assert np.array_equal(win.numpy(), g["window"])


@pytest.mark.parametrize("name", ["local_d3_even", "local_d4_odd", "local_d2_scaled"])
def test_local_naive_matches_reference(golden, name):
    g = golden(name)
    out = naive.local_match(g["prev_chw"].transpose(1, 2, 0), g["cur_chw"].transpose(1, 2, 0),
                            g["labels"], g["ids"].shape[0], int(g["d"]))
    assert np.max(np.abs(out - g["out"][0, :, :, :, 0])) <= 1e-5


def test_memory_session_matches_reference(golden):
    g = golden("memory_session")
    embs = torch.from_numpy(g["embs"])
    n_obj, d = int(g["n_obj"]), int(g["d"])
    T = embs.shape[0]
    gmem, lmem = {}, ({}, {})
    for rnd, ann in g["rounds"]:
        rnd, ann = int(rnd), int(ann)
        scr = torch.from_numpy(g[f"r{rnd}_scribble"])
        O.int_matching_step(embs[ann], scr, n_obj, d, gmem, lmem, "s", ann, rnd)
        for f in list(range(ann + 1, T)) + list(range(ann - 1, -1, -1)):
            prev_f = f - 1 if f > ann else f + 1
            pl = torch.from_numpy(g[f"r{rnd}_f{f}_prev_label"])
            gm, lm = O.prop_matching_step(embs[ann], embs[prev_f], embs[f], scr, pl, n_obj, 1, d, True,
                                          gmem, lmem, "s", f, rnd, ann)
            assert np.array_equal(gm[0, :, :, :, 0].permute(2, 0, 1).numpy(), g[f"r{rnd}_f{f}_global"])
            assert np.array_equal(lm[0, :, :, :, 0].permute(2, 0, 1).numpy(), g[f"r{rnd}_f{f}_local"])
    assert np.array_equal(gmem["s"][:T].numpy(), g["final_global_mem"])
    assert np.array_equal(lmem[0]["s"][:T, :3].numpy(), g["final_local_mem"])
    assert np.array_equal(lmem[1]["s"][:T, :3].numpy(), g["final_local_dist"])


def test_correlation_oracle_vs_naive_and_formula():
    gen = torch.Generator().manual_seed(5)
    a = torch.randn(2, 5, 7, 9, generator=gen)
    b = torch.randn(2, 5, 7, 9, generator=gen)
    for (pad, ks, md, s1, s2) in [(3, 1, 3, 1, 1), (4, 1, 4, 2, 2), (2, 3, 1, 1, 1), (0, 1, 0, 1, 1)]:
        got = O.correlation_forward(a, b, pad, ks, md, s1, s2).numpy()
        want = naive.correlation_forward(a.numpy(), b.numpy(), pad, ks, md, s1, s2)
        assert got.shape == want.shape
        assert np.max(np.abs(got - want)) < 1e-5
    # MANet's historical use (._bak/networks_old/IntVOS.py:264): pad=md=d, k=1, strides 1
    d = 2
    out = O.correlation_forward(a, b, d, 1, d, 1, 1)
    bp = torch.nn.functional.pad(b, (d, d, d, d))
    for tj in range(-d, d + 1):
        for ti in range(-d, d + 1):
            sh = bp[:, :, d + tj:d + tj + 7, d + ti:d + ti + 9]
            assert torch.allclose(out[:, (tj + d) * (2 * d + 1) + ti + d], (a * sh).mean(1), atol=1e-6)
    ga, gb = O.correlation_backward(a, b, torch.ones_like(out), d, 1, d, 1, 1)
    assert ga.shape == a.shape and gb.shape == b.shape


def test_normalize_is_exact_one_for_sentinel():
    x = torch.tensor([1e20, 0.0, 40.0])
    y = O.normalize_distance(x)
    assert y[0].item() == 1.0 and y[1].item() == 0.0 and y[2].item() == 1.0


# ------------------------------------------------------------------ gradients
def test_oracle_autograd_matches_reference_gradients(golden):
    """torch autograd through the oracle == autograd through the unmodified reference (golden vectors)."""
    g = golden("grad_global_k1")
    r = torch.from_numpy(g["ref_chw"]).requires_grad_(True)
    q = torch.from_numpy(g["query_chw"]).requires_grad_(True)
    out, _ = O.global_match(r.permute(1, 2, 0), q.permute(1, 2, 0), torch.from_numpy(g["labels"]).unsqueeze(-1), 1,
                            torch.tensor(int(g["n_obj"])), 5)
    (((torch.sigmoid(out) - 0.5) * 2) * torch.from_numpy(g["weights"])).sum().backward()
    assert np.array_equal(out.detach().numpy(), g["out"])
    assert np.abs(r.grad.numpy() - g["grad_ref_chw"]).max() <= 1e-6 * max(1.0, np.abs(g["grad_ref_chw"]).max())
    assert np.abs(q.grad.numpy() - g["grad_query_chw"]).max() <= 1e-6 * max(1.0, np.abs(g["grad_query_chw"]).max())

    g = golden("grad_local_d3")
    p = torch.from_numpy(g["prev_chw"]).requires_grad_(True)
    q = torch.from_numpy(g["cur_chw"]).requires_grad_(True)
    out = O.local_match(p.permute(1, 2, 0), q.permute(1, 2, 0), torch.from_numpy(g["labels"]).unsqueeze(-1),
                        torch.from_numpy(g["ids"]), int(g["d"]))
    (out * torch.from_numpy(g["weights"])).sum().backward()
    assert np.array_equal(out.detach().numpy(), g["out"])
    assert np.abs(p.grad.numpy() - g["grad_prev_chw"]).max() <= 1e-6 * max(1.0, np.abs(g["grad_prev_chw"]).max())
    assert np.abs(q.grad.numpy() - g["grad_query_chw"]).max() <= 1e-6 * max(1.0, np.abs(g["grad_query_chw"]).max())
