"""Generate the golden vectors under tests/golden/ by running the UNMODIFIED
reference (``/root/reference/networks/IntVOS.py``) on CPU through
``oracle/ref_shim.py``.  Run in the build container only:

    python tests/golden/make_golden.py

The reference has no tests or fixtures of its own for this path (SURVEY.md
section 4), so these files are the parity pin: seeded inputs + the reference's
outputs.  Everything is fp32/int32, small enough to keep in git.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import ref_shim  # noqa: E402


def emb(gen, c, h, w, dist):
    """[C,H,W] storage, the model's layout (IntVOS.py:605-606 permutes views of it)."""
    if dist == "A":   # the reference's own scratch-test distribution (._bak/ceshi.py:4-5)
        return torch.rand(c, h, w, generator=gen)
    return 0.1 * torch.relu(torch.randn(c, h, w, generator=gen))  # "B": post-BN-ReLU like


def blob_labels(gen, h, w, n_ids, cell=4):
    gh, gw = (h + cell - 1) // cell, (w + cell - 1) // cell
    grid = torch.randint(0, n_ids, (gh, gw), generator=gen)
    return grid.repeat_interleave(cell, 0).repeat_interleave(cell, 1)[:h, :w].int().contiguous()


def scribble_labels(gen, h, w, n_ids, absent=None, fill=-1):
    lab = torch.full((h, w), fill, dtype=torch.int32)
    for o in range(n_ids):
        if o == absent:
            continue
        y = int(torch.randint(0, h, (1,), generator=gen))
        x0 = int(torch.randint(0, max(1, w // 2), (1,), generator=gen))
        lab[y, x0:x0 + max(2, w // 3)] = o
        x = int(torch.randint(0, w, (1,), generator=gen))
        lab[max(0, y - 2):y + 3, x] = o
    return lab


def save(name, **arrays):
    out = {}
    for k, v in arrays.items():
        out[k] = v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print("wrote", name, {k: out[k].shape for k in out})


def global_cases():
    cases = [
        # name, seed, C, Hq, Wq, Hr, Wr, n_obj(gt_ids arg), k, dist, label kind, test_mode, n_chunks, pass_gt
        ("global_k1_A", 0, 100, 13, 17, 13, 17, 2, 1, "A", "blob", False, 10, True),
        ("global_k1_B_absent", 1, 100, 12, 16, 12, 16, 4, 1, "B", "blob_absent", False, 7, True),
        ("global_testmode_scribble", 2, 100, 14, 18, 14, 18, 3, 1, "B", "scribble", True, 10, True),
        ("global_k3_A", 3, 32, 10, 12, 10, 12, 2, 3, "A", "blob", False, 4, True),
        ("global_multiframe_ref", 4, 100, 9, 11, 27, 11, 2, 1, "B", "blob", False, 10, True),
        ("global_gtids_none", 5, 20, 8, 9, 8, 9, 3, 1, "A", "blob", False, 1, False),
        ("global_k2_testmode", 6, 24, 9, 10, 9, 10, 2, 2, "B", "scribble_dense", True, 3, True),
    ]
    for (name, seed, c, hq, wq, hr, wr, nobj, k, dist, kind, tm, nch, pass_gt) in cases:
        gen = torch.Generator().manual_seed(seed)
        ref_chw, qry_chw = emb(gen, c, hr, wr, dist), emb(gen, c, hq, wq, dist)
        n_ids = nobj + 1
        if kind == "blob":
            lab = blob_labels(gen, hr, wr, n_ids)
        elif kind == "blob_absent":
            lab = blob_labels(gen, hr, wr, n_ids)
            lab[lab == 2] = 0
        elif kind == "scribble":
            lab = scribble_labels(gen, hr, wr, n_ids, absent=2)
        else:
            lab = scribble_labels(gen, hr, wr, n_ids, absent=None)
            lab[::2, ::3] = torch.where(lab[::2, ::3] < 0, torch.zeros_like(lab[::2, ::3]), lab[::2, ::3])
        mod = ref_shim.load_reference(test_mode=tm)
        with ref_shim.cpu_cuda_identity(), torch.no_grad():
            out, ids = mod.nearest_neighbor_features_per_object(
                ref_chw.permute(1, 2, 0), qry_chw.permute(1, 2, 0), lab.unsqueeze(-1), k,
                torch.tensor(nobj) if pass_gt else None, n_chunks=nch)
        save(name, ref_chw=ref_chw, query_chw=qry_chw, labels=lab, out=out, ids=ids,
             k=k, n_obj=nobj, test_mode=int(tm), n_chunks=nch, pass_gt=int(pass_gt))


def selected_pixel_case():
    gen = torch.Generator().manual_seed(11)
    lab = scribble_labels(gen, 10, 13, 3).reshape(-1)
    e = torch.rand(lab.numel(), 8, generator=gen)
    mod = ref_shim.load_reference(test_mode=True)
    with ref_shim.cpu_cuda_identity():
        l2, e2 = mod._selected_pixel(lab, e)
    save("selected_pixel", labels=lab, emb=e, out_labels=l2, out_emb=e2)


def local_cases():
    cases = [
        ("local_d3_even", 20, 100, 12, 16, 2, 3, "B"),
        ("local_d4_odd", 21, 100, 13, 17, 3, 4, "B"),
        ("local_d12_window_gt_image", 22, 16, 26, 30, 2, 12, "A"),
        ("local_d9_A", 23, 100, 24, 28, 5, 9, "A"),
        ("local_d2_scaled", 24, 36, 14, 20, 2, 2, "B3"),
    ]
    for (name, seed, c, h, w, nobj, d, dist) in cases:
        gen = torch.Generator().manual_seed(seed)
        if dist == "B3":
            prev, cur = 3 * emb(gen, c, h, w, "B"), 3 * emb(gen, c, h, w, "B")
        else:
            prev = emb(gen, c, h, w, dist)
            # make the current frame a noisy copy so distances sit in the sensitive range
            cur = prev + 0.05 * torch.randn(c, h, w, generator=gen) if dist == "B" else emb(gen, c, h, w, dist)
        lab = blob_labels(gen, h, w, nobj + 1, cell=3)
        ids = torch.arange(0, nobj + 1).int()
        mod = ref_shim.load_reference()
        with ref_shim.cpu_cuda_identity(), torch.no_grad():
            out = mod.local_previous_frame_nearest_neighbor_features_per_object(
                prev.permute(1, 2, 0), cur.permute(1, 2, 0), lab.unsqueeze(-1), ids, max_distance=d)
            win = mod.local_pairwise_distances2(cur.permute(1, 2, 0), prev.permute(1, 2, 0), max_distance=d)
        extra = {"window": win} if d <= 4 else {}   # keep the fixtures small
        save(name, prev_chw=prev, cur_chw=cur, labels=lab, ids=ids, out=out, d=d, **extra)


def memory_session_case():
    """Drive the reference's own prop_seghead / int_seghead (IntVOS.py:583-764) for a
    3-round, 5-frame toy session and record the maps that reach the segmentation
    head (they are channels C and C+1 of its input) plus the final memories."""
    c, h, w, nobj, d, T = 12, 10, 12, 2, 3, 5
    gen = torch.Generator().manual_seed(77)
    embs = torch.stack([emb(gen, c, h, w, "B") for _ in range(T)])
    mod = ref_shim.load_reference(test_mode=True, max_local_distance=d)
    seen = []

    def fake_head(x):
        seen.append(x.detach().clone())
        return torch.zeros(x.shape[0], 1, x.shape[2], x.shape[3])

    fake_self = types.SimpleNamespace(inter_seghead=fake_head)
    gmem, lmem = {}, ({}, {})
    rounds = [(1, 2), (2, 0), (3, 3)]   # (interaction_num, annotated frame)
    log = {}
    prev_labels_store = {}
    with ref_shim.cpu_cuda_identity(), torch.no_grad():
        for (rnd, ann) in rounds:
            scr = scribble_labels(gen, h, w, nobj + 1, absent=(1 if rnd == 2 else None))
            log[f"r{rnd}_scribble"] = scr.clone()
            seen.clear()
            mod.IntVOS.int_seghead(fake_self, ref_frame_embedding=embs[ann:ann + 1],
                                   ref_scribble_label=scr.view(1, 1, h, w).float(), prev_round_label=None,
                                   global_map_tmp_dic=gmem, local_map_dics=lmem, interaction_num=rnd,
                                   seq_names=["s"], gt_ids=torch.tensor([nobj]), frame_num=[ann],
                                   first_inter=True)
            order = list(range(ann + 1, T)) + list(range(ann - 1, -1, -1))
            for f in order:
                prev_f = f - 1 if f > ann else f + 1
                pl = blob_labels(gen, h, w, nobj + 1, cell=3)
                prev_labels_store[(rnd, f)] = pl
                seen.clear()
                mod.IntVOS.prop_seghead(fake_self, ref_frame_embedding=embs[ann:ann + 1],
                                        previous_frame_embedding=embs[prev_f:prev_f + 1],
                                        current_frame_embedding=embs[f:f + 1],
                                        ref_scribble_label=scr.view(1, 1, h, w).float(),
                                        previous_frame_mask=pl.view(1, 1, h, w).float(),
                                        normalize_nearest_neighbor_distances=True, use_local_map=True,
                                        seq_names=["s"], gt_ids=torch.tensor([nobj]), k_nearest_neighbors=1,
                                        global_map_tmp_dic=gmem, local_map_dics=lmem, interaction_num=rnd,
                                        start_annotated_frame=ann, frame_num=[f], dynamic_seghead=fake_head)
                x = seen[-1]                       # [N, C+3, h, w]
                log[f"r{rnd}_f{f}_prev_label"] = pl
                log[f"r{rnd}_f{f}_global"] = x[:, c].clone()      # [N,h,w]
                log[f"r{rnd}_f{f}_local"] = x[:, c + 1].clone()
    save("memory_session", embs=embs, rounds=np.array(rounds), n_obj=nobj, d=d,
         final_global_mem=gmem["s"][:T], final_local_mem=lmem[0]["s"][:T, :3],
         final_local_dist=lmem[1]["s"][:T, :3], **log)


def gradient_cases():
    """Gradients of the UNMODIFIED reference (torch autograd through its own graph): the training use of the
    path (train_stage1.py:126).  loss = sum(w * normalised map) as the model consumes the maps."""
    # global, k = 1
    gen = torch.Generator().manual_seed(31)
    c, hr, wr, h, w, nobj = 32, 11, 13, 10, 12, 2
    ref = 0.3 * torch.randn(c, hr, wr, generator=gen)
    qry = 0.3 * torch.randn(c, h, w, generator=gen)
    lab = blob_labels(gen, hr, wr, nobj + 1)
    wts = torch.rand(1, h, w, nobj + 1, 1, generator=gen)
    mod = ref_shim.load_reference(test_mode=False)
    r, q = ref.clone().requires_grad_(True), qry.clone().requires_grad_(True)
    with ref_shim.cpu_cuda_identity():
        out, _ = mod.nearest_neighbor_features_per_object(r.permute(1, 2, 0), q.permute(1, 2, 0), lab.unsqueeze(-1), 1,
                                                          torch.tensor(nobj), n_chunks=5)
        (((torch.sigmoid(out) - 0.5) * 2) * wts).sum().backward()
    save("grad_global_k1", ref_chw=ref, query_chw=qry, labels=lab, weights=wts, n_obj=nobj, out=out,
         grad_ref_chw=r.grad, grad_query_chw=q.grad)
    # local
    gen = torch.Generator().manual_seed(32)
    c, h, w, nobj, d = 16, 14, 18, 2, 3
    prev = 0.3 * torch.randn(c, h, w, generator=gen)
    cur = prev + 0.15 * torch.randn(c, h, w, generator=gen)
    lab = blob_labels(gen, h, w, nobj + 1, cell=3)
    ids = torch.arange(0, nobj + 1).int()
    wts = torch.rand(1, h, w, nobj + 1, 1, generator=gen)
    p, q = prev.clone().requires_grad_(True), cur.clone().requires_grad_(True)
    with ref_shim.cpu_cuda_identity():
        out = mod.local_previous_frame_nearest_neighbor_features_per_object(p.permute(1, 2, 0), q.permute(1, 2, 0),
                                                                            lab.unsqueeze(-1), ids, max_distance=d)
        (out * wts).sum().backward()
    save("grad_local_d3", prev_chw=prev, cur_chw=cur, labels=lab, ids=ids, weights=wts, d=d, out=out,
         grad_prev_chw=p.grad, grad_query_chw=q.grad)


if __name__ == "__main__":
    assert ref_shim.reference_available(), "run this where /root/reference is mounted"
    torch.set_num_threads(4)
    global_cases()
    selected_pixel_case()
    local_cases()
    memory_session_case()
    gradient_cases()
