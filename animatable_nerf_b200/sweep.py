"""Frame-level drivers either side of `Renderer.render` (BASELINE config 5), all on the GPU:

 * `render_views`  -- a novel-view sweep: per view, rays + SMPL-box intersection + compaction on the device
   (`frontend.get_rays_within_bounds`, replacing the numpy code of `tpose_novel_view_dataset.__getitem__`,
   lib/datasets/tpose_novel_view_dataset.py:171-175), the fused render (with the silhouette culling of
   `tpose_renderer_mmsk` when the batch carries `msks`), and the scatter of the maps back into the (H, W) image
   through `mask_at_box` (what the visualizer does, lib/visualizers/if_nerf.py).  Views are the unit of multi-GPU
   sharding: view v goes to rank v % world.
 * `query_density_grid` -- the sigma cube of `aninerf_mesh_renderer.Renderer.render`
   (lib/networks/renderer/aninerf_mesh_renderer.py:26-44): `Network.calculate_alpha` over the `inside` points of a voxel
   grid in chunks of 2048*64 points; chunks are dealt round-robin to the ranks (the per-chunk argmin forcing of
   tpose_nerf_network.py:121 keeps the chunk a semantic unit) and gathered with one all_gather.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from . import _lib, frontend

GRID_CHUNK = 2048 * 64        # aninerf_mesh_renderer.py:35


def views_of_rank(n_views: int, rank: int, world: int):
    return list(range(rank, n_views, world))


@torch.no_grad()
def render_views(renderer, frame_batch: dict, K, w2c_list, H: int, W: int, rank: int = 0, world: int = 1, white_bkgd: bool = False):
    """frame_batch: the per-frame keys of a render batch (A, pbw, tbw, bounds, R, Th, latent indices [, msks, Ks, RT, H, W]) on the
    device; K (3,3); w2c_list: world->camera (4,4) per view.  Returns {view index: (rgb (H,W,3), acc (H,W), depth (H,W))} for the
    views of this rank, device tensors."""
    dev = frame_batch['A'].device
    wbounds = frame_batch['wbounds'].reshape(2, 3).cpu().numpy()
    out = {}
    for v in views_of_rank(len(w2c_list), rank, world):
        RT = np.asarray(w2c_list[v], dtype=np.float64)
        ray_o, ray_d, near, far, mask = frontend.get_rays_within_bounds(H, W, K, RT[:3, :3], RT[:3, 3:], wbounds, device=dev)
        rgb = torch.full((H * W, 3), 1.0 if white_bkgd else 0.0, device=dev)
        acc = torch.zeros(H * W, device=dev)
        depth = torch.zeros(H * W, device=dev)
        if ray_o.shape[0] > 0:
            b = dict(frame_batch)
            b['ray_o'], b['ray_d'], b['near'], b['far'] = ray_o[None], ray_d[None], near[None], far[None]
            r = renderer.render_device(b, want_bw=False)
            m = mask.view(-1)
            rgb[m] = r['rgb_map']
            acc[m] = r['acc_map']
            depth[m] = r['depth_map']
        out[v] = (rgb.view(H, W, 3), acc.view(H, W), depth.view(H, W))
    return out


def gather_views(local: dict, n_views: int, H: int, W: int, rank: int, world: int, device, group=None):
    """All ranks' rgb images -> (n_views, H, W, 3) on every rank: ONE all_gather of the (padded) per-rank stacks."""
    per = (n_views + world - 1) // world
    mine = torch.zeros(per, H, W, 3, device=device)
    for i, v in enumerate(views_of_rank(n_views, rank, world)):
        mine[i] = local[v][0]
    if world == 1:
        return mine[:n_views]
    recv = torch.empty(world * per, H, W, 3, device=device)
    dist.all_gather_into_tensor(recv, mine, group=group)
    order = [r * per + i for v in range(n_views) for r, i in [(v % world, v // world)]]
    return recv[torch.as_tensor(order, device=device)]


def grid_points(wbounds, voxel_size, device):
    """Voxel-centre-free grid of aninerf_mesh_dataset (lib/datasets/aninerf_mesh_dataset.py:141-152): arange(min, max + voxel, voxel)
    per axis, meshgrid 'ij' -> (X,Y,Z,3) float32."""
    wb = np.asarray(wbounds, dtype=np.float32).reshape(2, 3)
    vs = [float(v) for v in voxel_size]                      # cfg.voxel_size is a list of Python floats
    axes = [np.arange(wb[0, a], wb[1, a] + vs[a], vs[a]) for a in range(3)]
    pts = np.stack(np.meshgrid(*axes, indexing='ij'), axis=-1).astype(np.float32)
    return torch.from_numpy(pts).to(device)


def chunks_of_rank(n_pts: int, rank: int, world: int, chunk: int = GRID_CHUNK):
    return list(range(rank, (n_pts + chunk - 1) // chunk, world))


@torch.no_grad()
def query_density_grid(net, batch: dict, pts: torch.Tensor, inside: torch.Tensor = None, rank: int = 0, world: int = 1,
                       chunk: int = GRID_CHUNK, group=None):
    """pts (X,Y,Z,3) device, inside (X,Y,Z) bool or None (= all) -> sigma cube (X,Y,Z) float32 on every rank."""
    dev = pts.device
    shape = pts.shape[:-1]
    flat = pts.reshape(-1, 3)
    sel = None
    if inside is not None:
        sel = inside.reshape(-1).bool()
        flat = flat[sel]
    n = flat.shape[0]
    mine = chunks_of_rank(n, rank, world, chunk)
    idx = torch.cat([torch.arange(c * chunk, min(n, (c + 1) * chunk), device=dev) for c in mine]) if mine else torch.zeros(0, dtype=torch.long, device=dev)
    sigma_local = net.calculate_alpha(flat[idx].contiguous(), batch, chunk_pts=chunk) if idx.numel() else torch.zeros(0, device=dev)
    if world == 1:
        sigma = sigma_local
    else:
        n_chunks = (n + chunk - 1) // chunk
        per = ((n_chunks + world - 1) // world) * chunk
        send = torch.zeros(per, device=dev)
        send[:sigma_local.numel()] = sigma_local
        recv = torch.empty(world * per, device=dev)
        dist.all_gather_into_tensor(recv, send, group=group)
        pos = torch.empty(n, dtype=torch.long, device=dev)
        for r in range(world):
            cs = chunks_of_rank(n, r, world, chunk)
            if cs:
                ridx = torch.cat([torch.arange(c * chunk, min(n, (c + 1) * chunk), device=dev) for c in cs])
                pos[ridx] = r * per + torch.arange(ridx.numel(), device=dev)
        sigma = recv[pos]
    cube = torch.zeros(int(np.prod(shape)), device=dev)
    if sel is None:
        cube = sigma
    else:
        cube[sel] = sigma
    return cube.view(*shape)
