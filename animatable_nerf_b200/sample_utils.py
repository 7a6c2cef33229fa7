"""Drop-in for the one function of `lib/utils/sample_utils.py` that the extended networks' hot path calls
(`aligned_aninerf_lbw_network.py:65,98,112`, `aligned_aninerf_pdf_network.py:71,105`, `sdf_mesh_renderer.py:57,85`):
`sample_blend_closest_points` -- K-nearest SMPL vertices, inverse-distance weighted blend weights -- on the GPU through the C ABI
(`aninerf_knn_blend_weights`, csrc/train_ops.cu) instead of pytorch3d's `knn_points` + gathers + einsum.  (the new
piece those networks need; the networks themselves are not mirrored.)"""
from __future__ import annotations

import torch

from . import _lib


@torch.no_grad()
def sample_blend_closest_points(src: torch.Tensor, ref: torch.Tensor, values: torch.Tensor, K: int = 5, exp: float = 1e-8):
    """src (B,n,3), ref (B,V,3), values (B,V,24) -> sampled (B,n,24), dists (B,n,1)   (sample_utils.py:323-349)."""
    _lib.require_cuda(src, 'src')
    B, n, _ = src.shape
    sampled = torch.empty(B, n, 24, device=src.device)
    dists = torch.empty(B, n, 1, device=src.device)
    for b in range(B):
        p, v, w = _lib.f32c(src[b]), _lib.f32c(ref[b]), _lib.f32c(values[b].reshape(-1, 24))
        with torch.cuda.device(src.device):
            _lib.check(_lib.lib().aninerf_knn_blend_weights(_lib.ptr(p), n, _lib.ptr(v), v.shape[0], _lib.ptr(w), int(K), float(exp),
                                                            _lib.ptr(sampled[b]), _lib.ptr(dists[b]), _lib.stream_ptr(src.device)))
    return sampled, dists
