"""Drop-in `Renderer` for `lib/networks/renderer/aninerf_mesh_renderer.py` (select it with `renderer_module` / `renderer_path`):
the density cube of a posed / canonical body and its iso-surface.

Contract (aninerf_mesh_renderer.py:26-63): batch keys `pts (1,X,Y,Z,3)`, `inside (1,X,Y,Z)` (lib/datasets/aninerf_mesh_dataset.py
:157-166) plus the frame keys of the render batch; returns `{'vertex', 'posed_vertex', 'triangle'}` as numpy arrays -- vertices in
world coordinates `(v - 10) * cfg.voxel_size[0] + wbounds[0, 0]`, triangles as vertex-index triples.  Everything runs on the GPU:
`Network.calculate_alpha` over the inside points in chunks of 2048 * 64 (`sweep.query_density_grid`), zero-padding by 10, and
marching cubes at `cfg.mesh_th` through the C ABI (`aninerf_marching_cubes`, csrc/marching_cubes.cu) instead of PyMCubes on the
host.  The cube is also returned (`'cube'`, device tensor) for callers that want the raw densities.
"""
from __future__ import annotations

import torch

from . import _lib, config, sweep

PAD = 10          # aninerf_mesh_renderer.py:39


@torch.no_grad()
def marching_cubes(cube: torch.Tensor, iso: float):
    """cube (X,Y,Z) float32 CUDA tensor -> (vertices (V,3) float64 index coordinates, triangles (T,3) int32), device tensors.
    Same mesh as `mcubes.marching_cubes(cube, iso)` (vertex order: by owning grid point, then axis -- see oracle/marching_cubes.py)."""
    _lib.require_cuda(cube, 'cube')
    c = _lib.f32c(cube)
    X, Y, Z = c.shape
    L = _lib.lib()
    ws_bytes = L.aninerf_marching_cubes_workspace_bytes(X, Y, Z)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=c.device)
    counts = torch.zeros(2, dtype=torch.int32, device=c.device)
    cap_v, cap_t = 1 << 16, 1 << 17
    while True:
        verts = torch.empty(cap_v, 3, dtype=torch.float64, device=c.device)
        tris = torch.empty(cap_t, 3, dtype=torch.int32, device=c.device)
        with torch.cuda.device(c.device):
            _lib.check(L.aninerf_marching_cubes(_lib.ptr(c), X, Y, Z, float(iso), _lib.ptr(verts), cap_v, _lib.ptr(tris), cap_t, _lib.ptr(counts),
                                                _lib.ptr(ws), ws_bytes, _lib.stream_ptr(c.device)))
        nv, nt = (int(v) for v in counts.tolist())            # the mesh size is data dependent: one host read
        if nv <= cap_v and nt <= cap_t:
            return verts[:nv], tris[:nt]
        cap_v, cap_t = max(cap_v, nv), max(cap_t, nt)           # the counts are exact: the second pass fits


class Renderer:
    def __init__(self, net, cfg=None):
        self.net = net
        self.cfg = cfg if cfg is not None else getattr(net, 'cfg', None) or config.global_cfg()

    @torch.no_grad()
    def density_cube(self, batch, rank: int = 0, world: int = 1):
        """`cube[inside] = alpha` of aninerf_mesh_renderer.py:28-38 -> (X,Y,Z) float32 device tensor (chunks dealt to `world` ranks)."""
        pts = batch['pts']
        _lib.require_cuda(pts, "batch['pts']")
        inside = batch['inside'][0].bool() if 'inside' in batch else None
        return sweep.query_density_grid(self.net, batch, pts[0], inside, rank, world)

    @torch.no_grad()
    def render(self, batch):
        cfg = self.cfg
        cube = self.density_cube(batch)
        padded = torch.nn.functional.pad(cube, (PAD,) * 6)                       # np.pad(cube, 10, mode='constant')
        verts, tris = marching_cubes(padded, float(config.get(cfg, 'mesh_th')))
        voxel = float(config.get(cfg, 'voxel_size')[0])
        origin = batch['wbounds'].reshape(2, 3)[0].to(verts)
        vertices = ((verts - PAD) * voxel + origin).cpu().numpy()
        return {'vertex': vertices, 'posed_vertex': vertices, 'triangle': tris.cpu().numpy(), 'cube': cube}
