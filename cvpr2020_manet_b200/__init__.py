"""B200-native (sm_100a) implementation of MANet's pixel-embedding matching + map-memory hot
path, behind the reference's own Python API (include/manet_b200.h)."""
from .config import cfg  # noqa: F401

__all__ = ["cfg"]
