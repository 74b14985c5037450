"""Golden vectors for DynamicSegHead: runs the UNMODIFIED reference class
(``/root/reference/networks/IntVOS.py:510-525``) in eval mode on CPU through ``oracle/ref_shim.py``.
Build container only:

    python tests/golden/make_golden_seghead.py

The weights are the reference's own initialisation (kaiming-normal convs) with randomised batch-norm
affine parameters and running statistics, so that the BN folding is exercised.  fp16-compressible they
are not: the file is ~1.2 MB.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import ref_shim  # noqa: E402
from oracle import manet_oracle as O  # noqa: E402


def main():
    ref = ref_shim.load_reference()
    torch.manual_seed(7)
    head = ref.DynamicSegHead()
    gen = torch.Generator().manual_seed(11)
    with torch.no_grad():
        for name, mod in head.named_modules():
            if name.endswith("bn1") or name.endswith("bn2"):
                c = mod.weight.shape[0]
                mod.weight.copy_(0.5 + torch.rand(c, generator=gen))
                mod.bias.copy_(0.2 * torch.randn(c, generator=gen))
                mod.running_mean.copy_(0.1 * torch.randn(c, generator=gen))
                mod.running_var.copy_(0.5 + torch.rand(c, generator=gen))
    head.eval()
    n, c, h, w = 3, 100, 21, 37          # ragged against the 8x16 pixel tiles
    cur = 0.1 * torch.relu(torch.randn(c, h, w, generator=gen))
    gmap = torch.rand(1, h, w, n, 1, generator=gen)
    lmap = torch.rand(1, h, w, n, 1, generator=gen)
    lmap[0, :4, :, 1, 0] = 1.0
    prev = torch.randint(0, n, (h // 3, w // 3 + 1), generator=gen).repeat_interleave(3, 0).repeat_interleave(3, 1)[:h, :w].int()
    ids = torch.arange(n, dtype=torch.int32)
    x = O.seghead_features(cur, gmap, lmap, prev, ids)
    with torch.no_grad():
        y = head(x)
    state = {k: v.detach().clone() for k, v in head.state_dict().items()}
    mine = O.dynamic_seghead_forward(state, x)
    print("oracle vs reference max abs diff:", float((mine - y).abs().max()), "logit range", float(y.min()), float(y.max()))
    out = {"cur": cur.numpy(), "gmap": gmap.numpy(), "lmap": lmap.numpy(), "prev": prev.numpy(), "ids": ids.numpy(),
           "y": y.numpy()}                      # x = oracle.seghead_features(cur, gmap, lmap, prev, ids)
    for k in O.seghead_param_names():
        out["p:" + k] = state[k].numpy()
    np.savez_compressed(os.path.join(HERE, "seghead_ref.npz"), **out)
    print("wrote seghead_ref.npz", os.path.getsize(os.path.join(HERE, "seghead_ref.npz")) // 1024, "KB")
    prop_seghead_case(head)
    prop_seghead_grad_case(head)
    int_seghead_case()
    int_seghead_default_head_case()


def prop_seghead_case(head):
    """The reference's own IntVOS.prop_seghead (IntVOS.py:583-681) end to end -- matching, both memories and the
    head above -- for two propagation steps of one round; weights are those of seghead_ref.npz."""
    import types
    c, h, w, nobj, d = 100, 18, 22, 2, 3
    mod = ref_shim.load_reference(test_mode=True, max_local_distance=d)
    gen = torch.Generator().manual_seed(23)
    embs = torch.stack([0.1 * torch.relu(torch.randn(c, h, w, generator=gen)) for _ in range(3)])
    scr = torch.full((h, w), -1, dtype=torch.int32)
    scr[3, 2:12] = 0
    scr[9, 5:20] = 1
    scr[12:16, 7] = 2
    fake_self = types.SimpleNamespace()
    gmem, lmem = {}, ({}, {})
    out = {"embs": embs.numpy(), "scribble": scr.numpy(), "n_obj": nobj, "d": d}
    with ref_shim.cpu_cuda_identity(), torch.no_grad():
        for f in (1, 2):
            pl = torch.randint(0, nobj + 1, (h // 2, w // 2), generator=gen).repeat_interleave(2, 0).repeat_interleave(2, 1).int()
            res = mod.IntVOS.prop_seghead(fake_self, ref_frame_embedding=embs[0:1], previous_frame_embedding=embs[f - 1:f],
                                          current_frame_embedding=embs[f:f + 1], ref_scribble_label=scr.view(1, 1, h, w).float(),
                                          previous_frame_mask=pl.view(1, 1, h, w).float(),
                                          normalize_nearest_neighbor_distances=True, use_local_map=True, seq_names=["s"],
                                          gt_ids=torch.tensor([nobj]), k_nearest_neighbors=1, global_map_tmp_dic=gmem,
                                          local_map_dics=lmem, interaction_num=1, start_annotated_frame=0, frame_num=[f],
                                          dynamic_seghead=head)
            out[f"f{f}_prev_mask"] = pl.numpy()
            out[f"f{f}_pred"] = res[0]["s"].numpy()
    np.savez_compressed(os.path.join(HERE, "prop_seghead_ref.npz"), **out)
    print("wrote prop_seghead_ref.npz", {k: np.asarray(v).shape for k, v in out.items()})


def int_seghead_case():
    """The reference's own IntVOS.int_seghead (IntVOS.py:683-764) for two rounds on one sequence: the tensor its interaction
    head receives (captured by a stand-in head) and the memories it leaves behind."""
    import types
    c, h, w, nobj, d = 12, 14, 18, 2, 3
    mod = ref_shim.load_reference(test_mode=True, max_local_distance=d)
    gen = torch.Generator().manual_seed(41)
    embs = torch.stack([0.1 * torch.relu(torch.randn(c, h, w, generator=gen)) for _ in range(3)])
    seen = []

    def fake_head(x):
        seen.append(x.detach().clone())
        return torch.zeros(x.shape[0], 1, x.shape[2], x.shape[3])

    fake_self = types.SimpleNamespace(inter_seghead=fake_head)
    gmem, lmem = {}, ({}, {})
    out = {"embs": embs.numpy(), "n_obj": nobj, "d": d}
    with ref_shim.cpu_cuda_identity(), torch.no_grad():
        for rnd, frame, first in ((1, 1, True), (2, 2, False)):
            scr = torch.full((h, w), -1, dtype=torch.int32)
            scr[2 + rnd, 2:10] = 0
            scr[8, 4 + rnd:15] = 1
            scr[10:13, 6 + rnd] = 2
            prev_round = torch.randint(0, nobj + 1, (h // 2, w // 2), generator=gen).repeat_interleave(2, 0).repeat_interleave(2, 1).int()
            res = mod.IntVOS.int_seghead(fake_self, ref_frame_embedding=embs[frame:frame + 1],
                                         ref_scribble_label=scr.view(1, 1, h, w).float(),
                                         prev_round_label=None if first else prev_round.view(1, 1, h, w).float(),
                                         global_map_tmp_dic=gmem, local_map_dics=lmem, interaction_num=rnd, seq_names=["s"],
                                         gt_ids=torch.tensor([nobj]), frame_num=[frame], first_inter=first)
            out[f"r{rnd}_scribble"] = scr.numpy()
            out[f"r{rnd}_prev_round"] = prev_round.numpy()
            out[f"r{rnd}_to_cat"] = seen[-1].numpy()
            out[f"r{rnd}_pred_shape"] = np.array(res[0]["s"].shape)
    out["final_global_mem"] = gmem["s"][:3].numpy()
    out["final_local_dist"] = lmem[1]["s"][:3, :3].numpy()
    out["final_local_mem_is_ones"] = np.array(bool((lmem[0]["s"] == 1).all()))
    np.savez_compressed(os.path.join(HERE, "int_seghead_ref.npz"), **out)
    print("wrote int_seghead_ref.npz", {k: np.asarray(v).shape for k, v in out.items()})


def prop_seghead_grad_case(head):
    """Training form of IntVOS.prop_seghead (train_stage1.py:126: no memories, gradients into the embeddings): the
    reference's own autograd through both matchers, the normalisation and its torch DynamicSegHead (eval-mode BN, weights of
    seghead_ref.npz).  Records pred and d(sum(pred * weight))/d(embeddings)."""
    import types
    c, h, w, nobj, d = 100, 16, 20, 2, 3
    mod = ref_shim.load_reference(test_mode=False, max_local_distance=d)
    gen = torch.Generator().manual_seed(29)
    embs = torch.stack([0.1 * torch.relu(torch.randn(c, h, w, generator=gen)) + 0.01 * torch.randn(c, h, w, generator=gen)
                        for _ in range(3)]).requires_grad_(True)
    ref_lab = torch.randint(0, nobj + 1, (h // 4, w // 4), generator=gen).repeat_interleave(4, 0).repeat_interleave(4, 1).int()
    prev_lab = torch.randint(0, nobj + 1, (h // 2, w // 2), generator=gen).repeat_interleave(2, 0).repeat_interleave(2, 1).int()
    weight = torch.randn(1, nobj + 1, h, w, generator=gen)
    fake_self = types.SimpleNamespace()
    with ref_shim.cpu_cuda_identity():
        res = mod.IntVOS.prop_seghead(fake_self, ref_frame_embedding=embs[0:1], previous_frame_embedding=embs[1:2],
                                      current_frame_embedding=embs[2:3], ref_scribble_label=ref_lab.view(1, 1, h, w).float(),
                                      previous_frame_mask=prev_lab.view(1, 1, h, w).float(),
                                      normalize_nearest_neighbor_distances=True, use_local_map=True, seq_names=["s"],
                                      gt_ids=torch.tensor([nobj]), k_nearest_neighbors=1, global_map_tmp_dic=None,
                                      local_map_dics=None, interaction_num=None, start_annotated_frame=None, frame_num=None,
                                      dynamic_seghead=head)
        pred = res["s"]
        (pred * weight).sum().backward()
    out = {"embs": embs.detach().numpy(), "ref_label": ref_lab.numpy(), "prev_label": prev_lab.numpy(), "weight": weight.numpy(),
           "n_obj": nobj, "d": d, "pred": pred.detach().numpy(), "grad_embs": embs.grad.numpy()}
    np.savez_compressed(os.path.join(HERE, "prop_seghead_grad_ref.npz"), **out)
    print("wrote prop_seghead_grad_ref.npz", {k: np.asarray(v).shape for k, v in out.items()},
          "max |grad|", float(embs.grad.abs().max()))


def int_seghead_default_head_case():
    """IntVOS.int_seghead (IntVOS.py:683-764) with the interaction head of the reference's DEFAULT configuration:
    config.py:52 MODEL_USEIntSeg=False -> IntVOS.py:554 ``inter_seghead = DynamicSegHead(in_dim=C+2)``.  Two rounds (first
    interaction, then one with a previous-round label map); logits recorded."""
    import types
    c, h, w, nobj, d = 100, 18, 26, 3, 4
    mod = ref_shim.load_reference(test_mode=True, max_local_distance=d)
    torch.manual_seed(19)
    head = mod.DynamicSegHead(in_dim=c + 2)
    gen = torch.Generator().manual_seed(43)
    with torch.no_grad():
        for name, m in head.named_modules():
            if name.endswith("bn1") or name.endswith("bn2"):
                k = m.weight.shape[0]
                m.weight.copy_(0.5 + torch.rand(k, generator=gen))
                m.bias.copy_(0.2 * torch.randn(k, generator=gen))
                m.running_mean.copy_(0.1 * torch.randn(k, generator=gen))
                m.running_var.copy_(0.5 + torch.rand(k, generator=gen))
    head.eval()
    embs = torch.stack([0.1 * torch.relu(torch.randn(c, h, w, generator=gen)) for _ in range(3)])
    fake_self = types.SimpleNamespace(inter_seghead=head)
    gmem, lmem = {}, ({}, {})
    out = {"embs": embs.numpy(), "n_obj": nobj, "d": d}
    with ref_shim.cpu_cuda_identity(), torch.no_grad():
        for rnd, frame, first in ((1, 1, True), (2, 2, False)):
            scr = torch.full((h, w), -1, dtype=torch.int32)
            scr[2 + rnd, 2:12] = 0
            scr[8, 4 + rnd:19] = 1
            scr[10:15, 6 + rnd] = 2
            scr[15, 20:24] = 3
            prev_round = torch.randint(0, nobj + 1, (h // 2, w // 2), generator=gen).repeat_interleave(2, 0).repeat_interleave(2, 1).int()
            res = mod.IntVOS.int_seghead(fake_self, ref_frame_embedding=embs[frame:frame + 1],
                                         ref_scribble_label=scr.view(1, 1, h, w).float(),
                                         prev_round_label=None if first else prev_round.view(1, 1, h, w).float(),
                                         global_map_tmp_dic=gmem, local_map_dics=lmem, interaction_num=rnd, seq_names=["s"],
                                         gt_ids=torch.tensor([nobj]), frame_num=[frame], first_inter=first)
            out[f"r{rnd}_scribble"] = scr.numpy()
            out[f"r{rnd}_prev_round"] = prev_round.numpy()
            out[f"r{rnd}_pred"] = res[0]["s"].numpy()
    out["final_global_mem"] = gmem["s"][:3].numpy()
    for k, v in head.state_dict().items():
        if not k.endswith("num_batches_tracked"):
            out["p:" + k] = v.numpy()
    np.savez_compressed(os.path.join(HERE, "int_seghead_default_head_ref.npz"), **out)
    print("wrote int_seghead_default_head_ref.npz", os.path.getsize(os.path.join(HERE, "int_seghead_default_head_ref.npz")) // 1024, "KB",
          "logit range", float(out["r2_pred"].min()), float(out["r2_pred"].max()))


if __name__ == "__main__":
    main()
