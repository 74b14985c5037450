"""Host-side logic that needs no GPU: stride collapsing, shard arithmetic, flag table."""
import pytest
import torch

from cvpr2020_manet_b200 import distributed as D
from cvpr2020_manet_b200.config import cfg


def test_shard_bounds_partition_is_exact_and_balanced():
    for n in (0, 1, 7, 25680, 129600 * 8 + 3):
        for world in (1, 2, 3, 4, 8):
            spans = [D.shard_bounds(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        D.shard_bounds(10, 2, 2)


def test_config_defaults_match_reference_flags():
    # config.py:23,49,50,76 of the reference
    assert cfg.KNNS == 1 and cfg.MODEL_LOCAL_DOWNSAMPLE is True
    assert cfg.MODEL_MAX_LOCAL_DISTANCE == 12 and cfg.TEST_MODE is False


def test_pixel_view_collapses_permuted_chw_views_without_copy():
    from cvpr2020_manet_b200 import _device

    class FakeCuda(torch.Tensor):
        pass

    t = torch.rand(6, 4, 5)            # [C,H,W] storage
    v = t.permute(1, 2, 0)             # [H,W,C] view, strides (5,1,20)
    # bypass the device check: the stride arithmetic is what is under test
    orig = _device.require_f32
    _device.require_f32 = lambda *a, **k: None
    try:
        kept, p, c, ps, cs = _device.pixel_view(v, "v")
        assert (p, c, ps, cs) == (20, 6, 1, 20) and kept.data_ptr() == t.data_ptr()
        kept, p, c, ps, cs = _device.pixel_view(t.permute(2, 1, 0), "w")   # [W,H,C]: does not collapse
        assert (p, c, ps, cs) == (20, 6, 6, 1) and kept.is_contiguous()
        kept, p, c, ps, cs = _device.pixel_view(torch.rand(7, 3), "flat")
        assert (p, c, ps, cs) == (7, 3, 3, 1)
    finally:
        _device.require_f32 = orig
