"""world_size-2 gloo test of the reference-axis sharding logic (CPU).  The per-shard matcher is
the oracle here (the product matcher needs a GPU); what is under test is the host logic:
slice arithmetic, the MIN all-reduce and the post-reduction normalisation order."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from cvpr2020_manet_b200 import distributed as D
        from oracle import manet_oracle as O
        gen = torch.Generator().manual_seed(3)
        ref = torch.rand(9, 11, 16, generator=gen)
        qry = torch.rand(6, 7, 16, generator=gen)
        lab = torch.randint(0, 3, (9, 11, 1), generator=gen).int()
        lab[lab == 1] = 0 if rank >= 0 else 1      # object 1 absent everywhere
        lab[:5][lab[:5] == 2] = 0                  # object 2 only in the second shard

        def match(r, q, l, k, g, n_chunks):
            return O.global_match(r, q, l, k, g, n_chunks=1)

        got, ids = D.sharded_nearest_neighbor_features_per_object(ref, qry, lab, 1, torch.tensor(3), match_fn=match)
        want, _ = O.global_match(ref, qry, lab, 1, torch.tensor(3), n_chunks=1)
        ok = bool(torch.equal(got, want)) and got.shape == (1, 6, 7, 4, 1)
        ok = ok and bool((got[..., 1, 0] == 1e20).all()) and bool((got[..., 3, 0] == 1e20).all())
        ret[rank] = ok
    finally:
        dist.destroy_process_group()


def test_sharded_min_equals_unsharded_world2():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert all(ret.get(r) for r in range(world)), dict(ret)


def test_min_is_shard_order_invariant_numpy():
    rng = np.random.default_rng(0)
    d = rng.random((50, 400)).astype(np.float32)
    full = d.min(axis=1)
    for world in (2, 3, 8):
        parts = [d[:, b:e].min(axis=1) if e > b else np.full(50, 1e20, np.float32)
                 for b, e in [(r * 400 // world, (r + 1) * 400 // world) for r in range(world)]]
        assert np.array_equal(np.minimum.reduce(parts), full)
