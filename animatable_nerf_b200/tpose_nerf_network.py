"""Drop-in `Network` for `lib/networks/bw_deform/tpose_nerf_network.py` (select it with
`network_module` / `network_path`).

Same constructor, parameter names / shapes (so the reference's checkpoints and `load_network` work
unchanged) and the same public methods with the same tensor shapes -- but every method runs on the
sm_100a kernels of libaninerf_b200 through the C ABI.  There is no PyTorch fallback: tensors must be
on a CUDA device.

These methods are the inference entry points and run under `torch.no_grad()`; training (forward AND backward on the library's
kernels, gradients for every parameter of this module) goes through `tpose_trainer.TrainStep` -- `Renderer.render` under
`torch.enable_grad()` and `tpose_trainer.NetworkWrapper` -- not through these per-method calls.
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn as nn

from . import _lib, config
from .weights import PackedNet

_TRUNK_WIDTH = 256
_TRUNK_DEPTH = 8
_SKIP_AFTER = 4


def _trunk(input_ch: int) -> nn.ModuleList:
    """Eight 1x1 convs; the layer after index `_SKIP_AFTER` also sees the trunk input."""
    layers = []
    for i in range(_TRUNK_DEPTH):
        fan_in = input_ch if i == 0 else _TRUNK_WIDTH + (input_ch if i == _SKIP_AFTER + 1 else 0)
        layers.append(nn.Conv1d(fan_in, _TRUNK_WIDTH, 1))
    return nn.ModuleList(layers)


def _points(t: torch.Tensor) -> torch.Tensor:
    """(1,n,3) or (n,3) -> contiguous fp32 (n,3)"""
    t = t.reshape(-1, 3)
    return _lib.f32c(t)


def _frame_struct(batch: dict, need_tbw: bool = True):
    """Batch dict (tpose_dataset.py:236-277 schema) -> aninerf_frame; returns (struct, keep-alive list)."""
    keep = {}

    def dev(key):
        t = batch[key]
        _lib.require_cuda(t, f"batch['{key}']")
        t = _lib.f32c(t)
        keep[key] = t
        return t

    fr = _lib.Frame()
    fr.A = dev('A').data_ptr()
    fr.R = dev('R').data_ptr()
    fr.Th = dev('Th').data_ptr()
    fr.pbounds = dev('pbounds').data_ptr()
    fr.tbounds = dev('tbounds').data_ptr()
    pbw = dev('pbw')
    fr.pbw = pbw.data_ptr()
    fr.pbw_dims[:] = list(pbw.shape[-4:-1])
    if need_tbw and 'tbw' in batch:
        tbw = dev('tbw')
        fr.tbw = tbw.data_ptr()
        fr.tbw_dims[:] = list(tbw.shape[-4:-1])
    for key in ('latent_index', 'bw_latent_index'):
        if key not in batch:
            continue
        t = batch[key]
        if torch.is_tensor(t) and t.is_cuda and t.dtype == torch.int64:
            t = t.reshape(-1).contiguous()              # read by the kernels: no device->host sync per frame
            keep[key] = t
            setattr(fr, key + '_dev', t.data_ptr())
        else:
            setattr(fr, key, int(torch.as_tensor(t).reshape(-1)[0]))
    return fr, keep


class _KernelBacked:
    """Shared plumbing: lazily (re)pack the owning Network's parameters into an aninerf_net."""

    def _root(self):
        return self.__dict__['_root_ref']()


class TPoseHuman(nn.Module):
    """Canonical NeRF field; parameters as tpose_nerf_network.py:219-239."""

    def __init__(self, cfg):
        super().__init__()
        self.nf_latent = nn.Embedding(config.get(cfg, 'num_train_frame'), 128)
        self.actvn = nn.ReLU()
        self.pts_linears = _trunk(63)
        self.alpha_fc = nn.Conv1d(_TRUNK_WIDTH, 1, 1)
        self.feature_fc = nn.Conv1d(_TRUNK_WIDTH, _TRUNK_WIDTH, 1)
        self.latent_fc = nn.Conv1d(384, _TRUNK_WIDTH, 1)
        self.view_fc = nn.Conv1d(283, _TRUNK_WIDTH // 2, 1)
        self.rgb_fc = nn.Conv1d(_TRUNK_WIDTH // 2, 3, 1)
        self.__dict__['_owner'] = None

    @torch.no_grad()
    def calculate_alpha_rgb(self, nf_pts, viewdir, ind):
        """(1,m,3), (1,m,3), (1,) -> alpha (1,1,m), rgb (1,3,m), pre-activation (:252-275)."""
        net = self.__dict__['_owner']
        pts, vd = _points(nf_pts), _points(viewdir)
        m = pts.shape[0]
        sigma = torch.empty(m, device=pts.device)
        rgb = torch.empty(m, 3, device=pts.device)
        _lib.check(_lib.lib().aninerf_nerf_forward(net.packed().handle, int(ind.reshape(-1)[0]), _lib.ptr(pts), _lib.ptr(vd), m, None,
                                                   _lib.ptr(sigma), _lib.ptr(rgb), None, None, None, None, None,
                                                   net.nerf_precision, _lib.stream_ptr()))
        return sigma.view(1, 1, m), rgb.t().unsqueeze(0)

    @torch.no_grad()
    def calculate_alpha(self, nf_pts):
        """(1,m,3) -> alpha (1,1,m) (:241-250)."""
        pts = _points(nf_pts)
        ind = torch.zeros(1, dtype=torch.long)
        return self.calculate_alpha_rgb(nf_pts, pts.view(1, -1, 3), ind)[0]


class BackwardBlendWeight(nn.Module):
    """Novel-pose blend-weight field; parameters as tpose_nerf_network.py:279-294."""

    def __init__(self, cfg):
        super().__init__()
        self.bw_latent = nn.Embedding(config.get(cfg, 'num_eval_frame'), 128)
        self.actvn = nn.ReLU()
        self.bw_linears = _trunk(191)
        self.bw_fc = nn.Conv1d(_TRUNK_WIDTH, 24, 1)
        self.__dict__['_owner'] = None

    @torch.no_grad()
    def forward(self, ppts, smpl_bw, latent_index):
        return self.__dict__['_owner']._bw_field(_lib.FIELD_NOVEL_BW, ppts, smpl_bw, int(latent_index.reshape(-1)[0]))


class Network(nn.Module):
    def __init__(self, cfg=None):
        super().__init__()
        cfg = cfg if cfg is not None else config.global_cfg()
        self.__dict__['cfg'] = cfg
        # construction order follows the reference so that seeded default init yields the same values
        self.tpose_human = TPoseHuman(cfg)
        self.bw_latent = nn.Embedding(config.get(cfg, 'num_train_frame') + 1, 128)
        self.actvn = nn.ReLU()
        self.bw_linears = _trunk(191)
        self.bw_fc = nn.Conv1d(_TRUNK_WIDTH, 24, 1)
        if config.get(cfg, 'aninerf_animation'):
            self.novel_pose_bw = BackwardBlendWeight(cfg)
            self.novel_pose_bw.__dict__['_owner'] = self
        self.tpose_human.__dict__['_owner'] = self
        self.__dict__['_packed'] = None
        self.__dict__['_packed_key'] = None

    # ---------------------------------------------------------------------------------------
    @property
    def bw_precision(self):
        return int(config.get(self.cfg, 'b200_bw_precision'))

    @property
    def nerf_precision(self):
        return int(config.get(self.cfg, 'b200_nerf_precision'))

    def packed(self) -> PackedNet:
        """Kernel-side operand images, rebuilt whenever a parameter changed (in-place version bump,
        load_state_dict, .cuda())."""
        params = list(self.parameters())
        if not params or not params[0].is_cuda:
            raise _lib.AninerfError('Network parameters must be on a CUDA device (call .cuda()): no CPU fallback')
        key = tuple((p.data_ptr(), p._version) for p in params)
        if self.__dict__['_packed'] is None or self.__dict__['_packed_key'] != key:
            if self.__dict__['_packed'] is None:
                self.__dict__['_packed'] = PackedNet()
            self.__dict__['_packed'].load_state_dict(self.state_dict(), device=params[0].device)
            self.__dict__['_packed_key'] = key
        return self.__dict__['_packed']

    # ---------------------------------------------------------------------------------------
    @torch.no_grad()
    def _bw_field(self, field, pts, smpl_bw, latent_index, A=None):
        """pts (1,m,3), smpl_bw (1,24,m) -> bw (1,24,m) [, tpose (1,m,3) when A is given]"""
        p = _points(pts)
        m = p.shape[0]
        s = _lib.f32c(smpl_bw.reshape(24, m).t())                       # point-major (m,24)
        bw = torch.empty(m, 24, device=p.device)
        tp = torch.empty(m, 3, device=p.device) if A is not None else None
        Ac = _lib.f32c(A) if A is not None else None
        _lib.check(_lib.lib().aninerf_bw_forward(self.packed().handle, field, latent_index, _lib.ptr(p), _lib.ptr(s), m, None,
                                                 _lib.ptr(Ac), _lib.ptr(bw), _lib.ptr(tp), self.bw_precision, _lib.stream_ptr()))
        bw = bw.t().unsqueeze(0)
        return (bw, tp.unsqueeze(0)) if A is not None else bw

    def calculate_neural_blend_weights(self, pose_pts, smpl_bw, latent_index):
        """tpose_nerf_network.py:55-77."""
        return self._bw_field(_lib.FIELD_BW, pose_pts, smpl_bw, int(latent_index.reshape(-1)[0]))

    @torch.no_grad()
    def _sample_volume(self, pts, vol, bounds):
        """pts_sample_blend_weights (blend_utils.py:119-149): (1,m,3) -> (1,25,m)"""
        p = _points(pts)
        m = p.shape[0]
        v = _lib.f32c(vol)
        dims = (C.c_int32 * 3)(*v.shape[-4:-1])
        out = torch.empty(m, 25, device=p.device)
        b = _lib.f32c(bounds)                                            # (locals keep every operand alive across the launch)
        _lib.check(_lib.lib().aninerf_sample_blend_weights(_lib.ptr(p), m, _lib.ptr(v), dims, _lib.ptr(b), _lib.ptr(out),
                                                           _lib.stream_ptr()))
        return out.t().unsqueeze(0)

    @torch.no_grad()
    def pose_points_to_tpose_points(self, pose_pts, batch):
        """tpose_nerf_network.py:79-100 -> tpose (1,m,3), pbw (1,24,m)"""
        init_pbw = self._sample_volume(pose_pts, batch['pbw'], batch['pbounds'])[:, :24]
        if config.get(self.cfg, 'test_novel_pose'):
            field, idx = _lib.FIELD_NOVEL_BW, int(batch['bw_latent_index'].reshape(-1)[0])
        else:
            field, idx = _lib.FIELD_BW, int(batch['latent_index'].reshape(-1)[0]) + 1
        pbw, tpose = self._bw_field(field, pose_pts, init_pbw, idx, A=batch['A'])
        return tpose, pbw

    @torch.no_grad()
    def _world_to_pose(self, wpts, batch):
        w = _points(wpts)
        out = torch.empty_like(w)
        Rm, Th = _lib.f32c(batch['R']), _lib.f32c(batch['Th'])
        _lib.check(_lib.lib().aninerf_world_to_pose(_lib.ptr(w), w.shape[0], _lib.ptr(Rm), _lib.ptr(Th), _lib.ptr(out),
                                                    _lib.stream_ptr()))
        return out.unsqueeze(0)

    @torch.no_grad()
    def forward(self, wpts, viewdir, dists, batch):
        """Network.forward (tpose_nerf_network.py:139-215) for explicit sample points of ONE chunk:
        wpts/viewdir (n,3), dists (n,) -> {'pbw','tbw': (1,n'',24), 'raw': (1,n,4)}.
        (The renderer does not go through here: it uses the fused aninerf_render_rays path.)"""
        _lib.require_cuda(wpts, 'wpts')
        cfg = self.cfg
        n = wpts.shape[0]
        pose_pts = self._world_to_pose(wpts, batch)
        init_pbw = self._sample_volume(pose_pts, batch['pbw'], batch['pbounds'])
        pnorm = init_pbw[:, -1]
        pind = pnorm < config.get(cfg, 'norm_th')
        pind[torch.arange(len(pnorm)), pnorm.argmin(dim=1)] = True
        index = torch.nonzero(pind[0]).reshape(-1).to(torch.int32)
        sel = index.long()
        pose_sel = pose_pts[:, sel]
        vd = _lib.f32c(viewdir[sel])
        ds = _lib.f32c(dists[sel])
        tpose, pbw = self.pose_points_to_tpose_points(pose_sel, batch)
        init_tbw = self._sample_volume(tpose, batch['tbw'], batch['tbounds'])[:, :24]
        tbw = self._bw_field(_lib.FIELD_BW, tpose, init_tbw, 0)
        m = sel.numel()
        tp = _points(tpose)
        raw = torch.zeros(1, n, 4, device=wpts.device)
        sigma_masked = torch.empty(m, device=wpts.device)
        tb = _lib.f32c(batch['tbounds'])
        _lib.check(_lib.lib().aninerf_nerf_forward(self.packed().handle, int(batch['latent_index'].reshape(-1)[0]), _lib.ptr(tp),
                                                   _lib.ptr(vd), m, None, None, None, _lib.ptr(ds), _lib.ptr(tb), _lib.ptr(index),
                                                   _lib.ptr(raw), _lib.ptr(sigma_masked), self.nerf_precision, _lib.stream_ptr()))
        alpha_ind = sigma_masked > config.get(cfg, 'train_th')
        alpha_ind[torch.argmax(sigma_masked)] = True
        return {'pbw': pbw.transpose(1, 2)[:, alpha_ind], 'tbw': tbw.transpose(1, 2)[:, alpha_ind], 'raw': raw}

    @torch.no_grad()
    def calculate_alpha(self, wpts, batch, chunk_pts=None):
        """Network.calculate_alpha (tpose_nerf_network.py:105-137): wpts (m,3) -> sigma (m,), one chunk."""
        _lib.require_cuda(wpts, 'wpts')
        w = _points(wpts)
        m = w.shape[0]
        fr, keep = _frame_struct(batch, need_tbw=False)
        chunk = int(chunk_pts) if chunk_pts else max(256, (m + 255) // 256 * 256)
        pv = int(fr.pbw_dims[0]) * fr.pbw_dims[1] * fr.pbw_dims[2]
        ws_bytes = _lib.lib().aninerf_query_workspace_bytes(m, pv)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=w.device)
        sigma = torch.empty(m, device=w.device)
        n_active = torch.zeros(1, dtype=torch.int32, device=w.device)
        _lib.check(_lib.lib().aninerf_query_alpha(self.packed().handle, C.byref(fr), _lib.ptr(w), m, chunk, 0.1,
                                                  int(bool(config.get(self.cfg, 'test_novel_pose'))), self.bw_precision, _lib.ptr(sigma),
                                                  _lib.ptr(n_active), _lib.ptr(ws), ws_bytes, _lib.stream_ptr()))
        return sigma

    get_alpha = calculate_alpha
