"""The handful of flags the matching path reads, with the reference's defaults
(config.py:23,47,49,50,76 in the reference).  Unlike the reference, importing this module
neither parses ``sys.argv`` nor raises without CUDA; set attributes on ``cfg`` directly."""
from types import SimpleNamespace

cfg = SimpleNamespace(
    TEST_MODE=False,                    # --TEST_MODE, read at IntVOS.py:135,593
    MODEL_LOCAL_DOWNSAMPLE=True,        # read at IntVOS.py:225,279
    MODEL_MAX_LOCAL_DISTANCE=12,        # read at IntVOS.py:631,711
    MODEL_SEMANTIC_EMBEDDING_DIM=100,   # DynamicSegHead in_dim = this + 3 (IntVOS.py:511)
    MODEL_HEAD_EMBEDDING_DIM=256,       # DynamicSegHead width (IntVOS.py:511)
    TRAIN_BN_MOM=0.0003,
    KNNS=1,
)
