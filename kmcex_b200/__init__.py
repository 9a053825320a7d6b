"""kmcex_b200 -- B200 (sm_100a) build + retrieval path of a kmcEx model behind the reference's
KModel API.  Compute lives in libkmx.so (hand-written CUDA, C ABI in include/kmx.h)."""
from .kmodel import KModel, KmcDatabase, get_model  # noqa: F401
from ._lib import KmxError, lib  # noqa: F401
