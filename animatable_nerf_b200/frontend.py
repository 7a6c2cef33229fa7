"""GPU front end of a frame: ray generation, SMPL-box intersection and ray compaction, replacing the
numpy code of the reference's Dataset (`lib/utils/if_nerf/if_nerf_data_utils.py`:64-89, 156-196,
310-339).  Same names / argument meaning / return order; results are CUDA tensors.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib


def _camera(H, W, K, R, T):
    cam = _lib.Camera()
    cam.Kinv[:] = np.linalg.inv(np.asarray(K, dtype=np.float64)).ravel().tolist()   # host, as the reference (:83)
    cam.R[:] = np.asarray(R, dtype=np.float64).ravel().tolist()
    cam.T[:] = np.asarray(T, dtype=np.float64).ravel().tolist()
    cam.H, cam.W = int(H), int(W)
    return cam


@torch.no_grad()
def get_rays(H, W, K, R, T, device='cuda'):
    """-> rays_o, rays_d (H,W,3) float32 (the reference returns float64 and its callers cast, :328-329)."""
    cam = _camera(H, W, K, R, T)
    o = torch.empty(H * W, 3, device=device)
    d = torch.empty(H * W, 3, device=device)
    with torch.cuda.device(o.device):
        _lib.check(_lib.lib().aninerf_gen_rays(C.byref(cam), _lib.ptr(o), _lib.ptr(d), _lib.stream_ptr(o.device)))
    return o.view(H, W, 3), d.view(H, W, 3)


@torch.no_grad()
def get_near_far_dense(bounds, ray_o, ray_d):
    """near, far (n,) float32 for every ray (0 where the box is missed) and mask_at_box (n,) bool."""
    o, d = _lib.f32c(ray_o.reshape(-1, 3)), _lib.f32c(ray_d.reshape(-1, 3))
    n = o.shape[0]
    b = np.ascontiguousarray(np.asarray(bounds.cpu() if torch.is_tensor(bounds) else bounds, dtype=np.float32).reshape(6))
    near = torch.empty(n, device=o.device)
    far = torch.empty(n, device=o.device)
    mask = torch.empty(n, dtype=torch.uint8, device=o.device)
    with torch.cuda.device(o.device):
        _lib.check(_lib.lib().aninerf_near_far(b.ctypes.data_as(_lib.c_float_p), _lib.ptr(o), _lib.ptr(d), n, _lib.ptr(near),
                                               _lib.ptr(far), _lib.ptr(mask), _lib.stream_ptr(o.device)))
    return near, far, mask.bool()


@torch.no_grad()
def get_near_far(bounds, ray_o, ray_d):
    """get_near_far (:156-196): near, far only for the rays that hit the box, mask_at_box for all."""
    near, far, mask = get_near_far_dense(bounds, ray_o, ray_d)
    return near[mask], far[mask], mask


@torch.no_grad()
def get_rays_within_bounds(H, W, K, R, T, bounds, device='cuda'):
    """get_rays_within_bounds (:310-339) -> ray_o, ray_d (n,3), near, far (n,), mask_at_box (H,W) bool.
    One host sync (the ray count) per frame."""
    o, d = get_rays(H, W, K, R, T, device)
    o, d = o.view(-1, 3), d.view(-1, 3)
    n = o.shape[0]
    near, far, mask = get_near_far_dense(bounds, o, d)
    m8 = mask.to(torch.uint8)
    ws_bytes = _lib.lib().aninerf_compact_workspace_bytes(n)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=o.device)
    oo, do = torch.empty_like(o), torch.empty_like(d)
    no, fo = torch.empty_like(near), torch.empty_like(far)
    idx = torch.empty(n, dtype=torch.int32, device=o.device)
    cnt = torch.zeros(1, dtype=torch.int32, device=o.device)
    with torch.cuda.device(o.device):
        _lib.check(_lib.lib().aninerf_compact_rays(_lib.ptr(o), _lib.ptr(d), _lib.ptr(near), _lib.ptr(far), _lib.ptr(m8), n, _lib.ptr(oo),
                                                   _lib.ptr(do), _lib.ptr(no), _lib.ptr(fo), _lib.ptr(idx), _lib.ptr(cnt), _lib.ptr(ws),
                                                   ws_bytes, _lib.stream_ptr(o.device)))
    k = int(cnt.item())
    return oo[:k], do[:k], no[:k], fo[:k], mask.view(H, W)
