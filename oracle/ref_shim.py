"""Import shim for the UNMODIFIED reference hot path (test infrastructure only).

This file is part of the ORACLE side of the repo: it may be imported only by
``tests/``, ``tests/golden/make_golden.py`` and by nothing on the product path.
It works only where ``/root/reference`` is mounted (the build container); the
GPU box never sees it -- the golden vectors it produced travel instead.

Why a shim is needed (reference facts, nothing is copied):
  * ``config.py:79-83`` parses ``sys.argv`` at import and raises without CUDA,
    so a stub ``config`` module carrying the handful of flags the hot path reads
    (``IntVOS.py:135,225,279,593``) is placed in ``sys.modules`` first.
  * ``IntVOS.py:102,200`` call ``.cuda()`` unconditionally; on a CPU-only host we
    temporarily make ``Tensor.cuda`` the identity while the reference runs.
"""
from __future__ import annotations

import contextlib
import importlib
import os
import sys
import types

import torch

REFERENCE_ROOT = os.environ.get("MANET_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "networks", "IntVOS.py"))


def _stub_config(test_mode: bool, max_local_distance: int) -> types.ModuleType:
    cfg = types.SimpleNamespace(
        TEST_MODE=test_mode,
        MODEL_LOCAL_DOWNSAMPLE=True,
        MODEL_MAX_LOCAL_DISTANCE=max_local_distance,
        MODEL_SEMANTIC_EMBEDDING_DIM=100,
        MODEL_HEAD_EMBEDDING_DIM=256,
        MODEL_ASPP_OUTDIM=256,
        TRAIN_BN_MOM=0.0003,
        MODEL_USEIntSeg=False,
        KNNS=1,
    )
    mod = types.ModuleType("config")
    mod.cfg = cfg
    return mod


_cached = {}


def load_reference(test_mode: bool = False, max_local_distance: int = 12):
    """Return the reference ``networks.IntVOS`` module (imported once) with its
    module-level ``cfg`` flags set as requested."""
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    if "mod" not in _cached:
        sys.modules["config"] = _stub_config(test_mode, max_local_distance)
        if REFERENCE_ROOT not in sys.path:
            sys.path.insert(0, REFERENCE_ROOT)
        _cached["mod"] = importlib.import_module("networks.IntVOS")
    mod = _cached["mod"]
    mod.cfg.TEST_MODE = test_mode
    mod.cfg.MODEL_MAX_LOCAL_DISTANCE = max_local_distance
    return mod


@contextlib.contextmanager
def cpu_cuda_identity():
    """Make ``Tensor.cuda()`` a no-op while the reference runs on a CPU host."""
    if torch.cuda.is_available():
        yield
        return
    saved = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self  # type: ignore[assignment]
    try:
        yield
    finally:
        torch.Tensor.cuda = saved  # type: ignore[assignment]
