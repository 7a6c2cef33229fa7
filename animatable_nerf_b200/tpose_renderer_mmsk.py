"""Drop-in `Renderer` for `lib/networks/renderer/tpose_renderer_mmsk.py` (novel-view / pose-sequence
rendering with multi-view silhouette culling; select it with `renderer_module` / `renderer_path`).

Contract (tpose_renderer_mmsk.py:99-166): the `tpose_renderer` batch plus `msks (1,V,H,W) u8`,
`Ks (1,V,3,3)`, `RT (1,V,4,4)`, `H`, `W` (lib/datasets/tpose_novel_view_dataset.py:191) in;
`rgb_map (1,R,3)`, `acc_map (1,R)`, `depth_map (1,R)` out, always detached CPU tensors.  Samples that
do not project into every training-view silhouette never reach the network: the culling runs inside the
front-end mask kernel of the fused path (csrc/geometry.cu `inside_all_views`), so the per-chunk
argmin forcing of `Network.forward` sees the survivors only, as in the reference.
"""
from __future__ import annotations

import torch

from . import _lib, config, tpose_renderer


def silhouettes_struct(batch):
    """(_lib.Silhouettes, keep-alive tensors) from the batch keys of tpose_novel_view_dataset.py:191."""
    if 'Ks' not in batch or 'msks' not in batch or 'RT' not in batch:
        raise KeyError("tpose_renderer_mmsk needs batch['msks'], batch['Ks'], batch['RT'], batch['H'], batch['W']")
    msks = batch['msks']
    _lib.require_cuda(msks, "batch['msks']")
    msks = msks[0] if msks.dim() == 4 else msks
    if msks.dtype == torch.bool:
        msks = msks.to(torch.uint8)
    if msks.dtype != torch.uint8:
        msks = (msks != 0).to(torch.uint8)
    msks = msks.contiguous()
    dev = msks.device
    Ks = _lib.f32c(batch['Ks'].to(dev).reshape(-1, 3, 3))
    RT = _lib.f32c(batch['RT'].to(dev).reshape(-1, 4, 4))
    V, H, W = msks.shape
    # the reference clamps with batch['H'], batch['W'] (tpose_renderer_mmsk.py:43-45) and indexes the (H,W) mask
    bh, bw = int(batch['H']) if 'H' in batch else H, int(batch['W']) if 'W' in batch else W
    if (bh, bw) != (H, W):
        raise _lib.AninerfError(f"batch['H'], batch['W'] = {(bh, bw)} do not match the mask planes {(H, W)}")
    if Ks.shape[0] != V or RT.shape[0] != V:
        raise _lib.AninerfError('msks, Ks and RT disagree on the number of views')
    s = _lib.Silhouettes(msks=msks.data_ptr(), Ks=Ks.data_ptr(), RT=RT.data_ptr(), n_views=V, H=H, W=W)
    return s, (msks, Ks, RT)


class Renderer(tpose_renderer.Renderer):
    def __init__(self, net, cfg=None):
        super().__init__(net, cfg)

    @torch.no_grad()
    def prepare_inside_pts(self, pts, batch):
        """pts (1, chunk, S, 3) world points -> bool (1, chunk*S): inside every training-view silhouette
        (tpose_renderer_mmsk.py:14-57)."""
        _lib.require_cuda(pts, 'pts')
        p = _lib.f32c(pts.reshape(-1, 3))
        sil, keep = silhouettes_struct(batch)
        out = torch.empty(p.shape[0], dtype=torch.uint8, device=p.device)
        import ctypes as C
        _lib.check(_lib.lib().aninerf_inside_all_views(_lib.ptr(p), p.shape[0], C.byref(sil), _lib.ptr(out), _lib.stream_ptr(p.device)))
        return out.bool().view(pts.shape[0], -1)

    @torch.no_grad()
    def render_device(self, batch, t_rand=None, want_bw=None, silhouettes=None, peers=None, keep_raw=False):
        """Same parameters as tpose_renderer.Renderer.render_device (incl. `peers`: the fused peer-memory image gather of the
        ray-tiled multi-GPU path); the culled renderer never produces pbw / tbw (tpose_renderer_mmsk.py:135-139)."""
        if silhouettes is None:
            silhouettes = silhouettes_struct(batch)
        return super().render_device(batch, t_rand=t_rand, want_bw=False, silhouettes=silhouettes, peers=peers, keep_raw=keep_raw)

    @torch.no_grad()
    def render(self, batch):
        cfg = self.cfg
        ray_o = batch['ray_o']
        R = ray_o.shape[1]
        S = int(config.get(cfg, 'N_samples'))
        t_rand = None
        if config.get(cfg, 'perturb') > 0. and self.net.training:
            t_rand = torch.rand(1, R, S).to(ray_o.device)
        out = self.render_device(batch, t_rand=t_rand)
        ret = {'rgb_map': out['rgb_map'].view(1, R, 3), 'acc_map': out['acc_map'].view(1, R), 'depth_map': out['depth_map'].view(1, R)}
        # tpose_renderer_mmsk.py:135-139: always detached host tensors
        return self._to_host(ret, ray_o.device)
