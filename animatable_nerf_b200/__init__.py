"""animatable_nerf_b200 -- Animatable NeRF's per-ray render hot path on B200 (sm_100a).

Host side: Python mirrors of the reference's `Renderer` / `Network` / ray-generation interface.
Device side: libaninerf_b200.so (hand-written CUDA incl. tcgen05 MLP kernels) behind the C ABI of
include/aninerf_b200.h.  No CPU fallback, no PyTorch compute fallback.
"""
from . import config  # noqa: F401
from ._lib import AninerfError, LIB_PATH  # noqa: F401

__all__ = ['config', 'AninerfError', 'LIB_PATH']
