"""Drop-in for `lib/train/trainers/aninerf_animation_trainer.py` (the second training stage: fit `net.novel_pose_bw`, the
blend-weight field of unseen poses, against the frozen stage-1 fields; no rendering).

`NetworkWrapper(net).forward(batch)` -> `(ret, loss, scalar_stats, image_stats)` with
`loss = smooth_l1(pbw0, tbw0) + smooth_l1(pbw1, tbw1)` (:33-53): 65 536 uniform samples of `wbounds` go
observation -> canonical (`ppts_to_tpose`, :56-91) and 65 536 samples of `tbounds` go canonical -> observation
(`tpose_to_ppts`, :94-119).  Only `novel_pose_bw.*` receives gradients (:26-31); the frozen canonical field still
back-propagates DATA gradients (tbw depends on the canonical point, which depends on the trained field through the
inverse LBS).  Forward and backward run on the kernels of the training step (csrc/gemm_x3.cu, csrc/train_ops.cu).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib, config
from . import train_ops as T
from .train_ops import Op
from .tpose_trainer import _Grads, _PlannedTrunk, _w2

N_POINTS = 1024 * 64      # get_sampling_points, :131


def get_sampling_points(bounds, n=N_POINTS):
    """(1,2,3) bounds -> (1,n,3) uniform samples; torch.rand on the CPU generator, x then y then z (:122-142)."""
    lo, hi = bounds[:, 0], bounds[:, 1]
    vals = torch.stack([torch.rand([bounds.shape[0], n]) for _ in range(3)], dim=2).to(bounds.device)
    return (hi - lo)[:, None] * vals + lo[:, None]


class AnimationStep:
    def __init__(self, net, cfg=None):
        self.net = net
        self.cfg = cfg if cfg is not None else getattr(net, 'cfg', None) or config.global_cfg()
        self._plan = None

    def _ensure_plan(self, m0, m1, dev, sd):
        """Flat gradient buffer + the six trunk evaluations of a step as planned trunks (static buffers, prebuilt descriptors):
        observation->canonical: novel field, frozen canonical field (data gradients only), frozen density trunk;
        canonical->observation: frozen canonical field, frozen density trunk, novel field."""
        key = (m0, m1, str(dev), tuple(p.data_ptr() for p in sd.values()))
        if self._plan is None or self._plan['key'] != key:
            G = _Grads(self.net)
            d_pe = torch.zeros(m0, 64, device=dev)
            nv, bw, nf = 'novel_pose_bw.bw_linears', 'bw_linears', 'tpose_human.pts_linears'
            p0 = _PlannedTrunk(sd, nv, 128, G, m0, dev)
            c0 = _PlannedTrunk(sd, bw, 128, G, m0, dev, d_pe=d_pe)
            n0 = _PlannedTrunk(sd, nf, 0, G, m0, dev, pe=c0.pe)
            c1 = _PlannedTrunk(sd, bw, 128, G, m1, dev)
            n1 = _PlannedTrunk(sd, nf, 0, G, m1, dev, pe=c1.pe)
            q1 = _PlannedTrunk(sd, nv, 128, G, m1, dev)
            self._plan = {'key': key, 'G': G, 'd_pe': d_pe, 'p0': p0, 'c0': c0, 'n0': n0, 'c1': c1, 'n1': n1, 'q1': q1}
        return self._plan

    @torch.no_grad()
    def run(self, batch, wpts, tpts):
        """wpts, tpts (n,3) device.  Returns (ret, stats, grads)."""
        cfg, net = self.cfg, self.net
        sd = dict(net.named_parameters())
        if 'novel_pose_bw.bw_fc.weight' not in sd:
            raise _lib.AninerfError('the network has no novel_pose_bw field (cfg.aninerf_animation must be True)')
        _lib.require_cuda(wpts, 'wpts')
        dev = wpts.device
        T.begin_step(dev)
        wpts, tpts = _lib.f32c(wpts.reshape(-1, 3)), _lib.f32c(tpts.reshape(-1, 3))
        plan = self._ensure_plan(wpts.shape[0], tpts.shape[0], dev, sd)
        G = plan['G']
        G.flat.zero_()
        A = _lib.f32c(batch['A'].reshape(24, 4, 4))
        pvol, tvol = _lib.f32c(batch['pbw'][0]), _lib.f32c(batch['tbw'][0])
        pb, tb = _lib.f32c(batch['pbounds'].reshape(2, 3)), _lib.f32c(batch['tbounds'].reshape(2, 3))
        Rm, Th = _lib.f32c(batch['R'].reshape(3, 3)), _lib.f32c(batch['Th'].reshape(3))
        idx = int(torch.as_tensor(batch['bw_latent_index']).reshape(-1)[0])
        lat_np = sd['novel_pose_bw.bw_latent.weight'].detach()[idx:idx + 1]
        g_lat_np = G.views['novel_pose_bw.bw_latent.weight'][idx:idx + 1]
        lat0 = sd['bw_latent.weight'].detach()[0:1]
        Wfc_np, bfc_np = _w2(sd['novel_pose_bw.bw_fc.weight']), sd['novel_pose_bw.bw_fc.bias'].detach()
        gWfc_np, gbfc_np = G.w('novel_pose_bw.bw_fc.weight'), G.views['novel_pose_bw.bw_fc.bias']
        Wfc, bfc = _w2(sd['bw_fc.weight']), sd['bw_fc.bias'].detach()
        Wa, ba = _w2(sd['tpose_human.alpha_fc.weight']), sd['tpose_human.alpha_fc.bias'].detach()
        norm_th, train_th = float(config.get(cfg, 'norm_th')), float(config.get(cfg, 'train_th'))

        def e(*shape):
            return torch.empty(*shape, device=dev)

        def novel_field(trunk, pts, vol, bounds):
            """novel_pose_bw(pts, init, idx) on a planned trunk: returns (h8, init25, bw)"""
            m = pts.shape[0]
            init = T.sample_volume(pts, vol, bounds, e(m, 25))
            T.pe_forward(pts, 10, trunk.pe[:m])
            h8 = trunk.forward(m, lat_np)
            delta = T.gemm([(Op(h8), Op(Wfc_np))], e(m, 24), bias=bfc_np)
            return h8, init, T.bw_softmax_forward(init, delta, e(m, 24))

        def canonical_field(trunk, pts):
            """net.calculate_neural_blend_weights(pts, init_tbw, 0) with the frozen stage-1 weights"""
            m = pts.shape[0]
            init = T.sample_volume(pts, tvol, tb, e(m, 25))
            T.pe_forward(pts, 10, trunk.pe[:m])
            h8 = trunk.forward(m, lat0)
            delta = T.gemm([(Op(h8), Op(Wfc))], e(m, 24), bias=bfc)
            return h8, init, T.bw_softmax_forward(init, delta, e(m, 24))

        def density(trunk, m):
            """net.tpose_human.calculate_alpha(pts) (frozen); the trunk shares the PE buffer of the canonical field"""
            h8 = trunk.forward(m, None)
            return T.gemm([(Op(h8), Op(Wa))], e(m, 1), bias=ba)

        def select(sigma_masked):
            m = sigma_masked.shape[0]
            sel = torch.empty(m, dtype=torch.uint8, device=dev)
            n_sel = torch.zeros(1, dtype=torch.int32, device=dev)
            offs = torch.tensor([0, m], dtype=torch.int32, device=dev)
            T.select_rows(sigma_masked, offs, 1, train_th, sel, n_sel)
            return sel, n_sel

        def novel_backward(trunk, h8, init, bw, d_bw):
            m = bw.shape[0]
            d_delta = e(m, 24)
            T.bw_softmax_backward(init, bw, d_bw, d_delta, None)
            T.gemm([(Op(d_delta).T, Op(h8).T)], gWfc_np, accumulate=True, split_k=T.split_for(m))
            T.colsum(d_delta, gbfc_np, accumulate=True)
            T.gemm([(Op(d_delta), Op(Wfc_np).T)], trunk.dz[0][:m], relu_mask=h8)
            trunk.backward(m, g_lat_np, False)

        losses = torch.zeros(2, device=dev)
        # ---- observation -> canonical (ppts_to_tpose) ---------------------------------------------------------------------
        m0 = wpts.shape[0]
        ppts = T.world_to_pose(wpts, Rm, Th, e(m0, 3))
        trunk_p, trunk_c = plan['p0'], plan['c0']
        h8p, init_p, pbw = novel_field(trunk_p, ppts, pvol, pb)
        tpose = T.inverse_lbs(ppts, pbw, A, e(m0, 3))
        h8c, init_t, tbw = canonical_field(trunk_c, tpose)
        sigma = density(plan['n0'], m0)
        sm = T.mask_sigma(sigma, tpose, tb, init_p[:, 24:], 25, norm_th, e(m0))
        sel0, n_sel0 = select(sm)
        d_pbw, d_tbw = e(m0, 24), e(m0, 24)
        T.bw_loss(pbw, tbw, sel0, n_sel0, losses[0:1], d_pbw, d_tbw)
        # ---- canonical -> observation (tpose_to_ppts) ---------------------------------------------------------------------
        m1 = tpts.shape[0]
        _, _, tbw_c = canonical_field(plan['c1'], tpts)
        sigma1 = density(plan['n1'], m1).view(-1)
        sel1, n_sel1 = select(sigma1)
        pose_pts = T.forward_lbs(tpts, tbw_c, A, e(m1, 3))
        trunk_q = plan['q1']
        h8q, init_q, pbw_c = novel_field(trunk_q, pose_pts, pvol, pb)
        d_pbw_c, d_unused = e(m1, 24), e(m1, 24)
        T.bw_loss(pbw_c, tbw_c, sel1, n_sel1, losses[1:2], d_pbw_c, d_unused)

        # ================================= backward ========================================================================
        novel_backward(trunk_q, h8q, init_q, pbw_c, d_pbw_c)
        # frozen canonical field: data gradients only, down to the canonical point
        d_delta, d_init = e(m0, 24), e(m0, 24)
        T.bw_softmax_backward(init_t, tbw, d_tbw, d_delta, d_init)
        T.gemm([(Op(d_delta), Op(Wfc).T)], trunk_c.dz[0][:m0], relu_mask=h8c)
        d_pe = plan['d_pe'][:m0]
        trunk_c.backward(m0, None, True, wgrad=False)
        d_tp = T.pe_backward(tpose, d_pe, 10, e(m0, 3), False)
        T.sample_volume_backward(tpose, tvol, tb, d_init, d_tp, True)
        T.inverse_lbs_backward(pbw, A, tpose, d_tp, d_pbw, True)
        novel_backward(trunk_p, h8p, init_p, pbw, d_pbw)

        ret = {'pbw0': pbw[sel0.bool()]}
        stats = {'bw_loss0': losses[0], 'bw_loss1': losses[1], 'loss': losses[0] + losses[1]}
        self._keep = (A, pvol, tvol, pb, tb, Rm, Th, wpts, tpts)
        T.end_step(dev)
        return ret, stats, G


class _LossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, loss, grads, *params):
        ctx.grads = grads
        return loss.clone()

    @staticmethod
    def backward(ctx, g):
        return (None, None, *[gr * g for gr in ctx.grads])


class NetworkWrapper(nn.Module):
    def __init__(self, net, cfg=None):
        super().__init__()
        self.net = net
        self.cfg = cfg if cfg is not None else getattr(net, 'cfg', None) or config.global_cfg()
        for p in self.net.parameters():                       # :26-31
            p.requires_grad = False
        for p in self.net.novel_pose_bw.parameters():
            p.requires_grad = True
        self.__dict__['_step'] = AnimationStep(net, self.cfg)

    def forward(self, batch, wpts=None, tpts=None):
        dev = batch['A'].device
        if wpts is None:
            wpts = get_sampling_points(batch['wbounds'])[0]
        if tpts is None:
            tpts = get_sampling_points(batch['tbounds'])[0]
        ret, stats, G = self.__dict__['_step'].run(batch, wpts.to(dev), tpts.to(dev))
        names = [k for k, p in self.net.named_parameters() if p.requires_grad]
        params = [p for _, p in self.net.named_parameters() if p.requires_grad]
        loss = _LossFn.apply(stats['loss'], [G.views[k] for k in names], *params)
        return ret, loss, {'bw_loss0': stats['bw_loss0'], 'bw_loss1': stats['bw_loss1'], 'loss': loss}, {}
