"""Drop-in for `lib/train/trainers/tpose_trainer.py`: `NetworkWrapper(net).forward(batch)` returns
`(ret, loss, scalar_stats, image_stats)` with `loss = smooth_l1(pbw, tbw) + mse(rgb_map[mask], rgb[mask])`
(tpose_trainer.py:21-73), and `loss.backward()` fills the `.grad` of every `Network` parameter exactly as
`Trainer.train` expects (lib/train/trainers/trainer.py:62-66) -- but forward AND backward run on this library's
kernels: the front end of the fused render path, then layer-by-layer split-precision tcgen05 GEMMs
(csrc/gemm_x3.cu: forward, data gradient and weight gradient of every 1x1 Conv1d, fp32-equivalent) and the
per-point forward/backward kernels of csrc/train_ops.cu.  Activations are kept in fp32, so losses and gradients
match the fp32 reference to ~1e-5 relative.  PyTorch only owns the memory, the parameter / optimizer objects
(torch.optim.Adam as in lib/train/optimizer.py:12-27) and NCCL (`allreduce_gradients`).

Gradient flow (what torch.autograd derives for tpose_nerf_network.py:139-215):
  img loss -> raw2outputs -> tail -> NeRF heads/trunk -> PE(tpose) -----------.
  bw loss  -> tbw -> softmax -> canonical blend-weight trunk -> PE(tpose) ----+--> d tpose -> inverse LBS -> d pbw
           -> tbw -> log(init_tbw) -> trilinear sampling at tpose ------------'                               |
  bw loss  -> pbw ---------------------------------------------------------------------------------------------+--> softmax -> posed trunk
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn as nn

from . import _lib, config
from . import train_ops as T
from .train_ops import Op
from .tpose_nerf_network import _frame_struct
from .tpose_renderer import Renderer


def _w2(p):
    """Conv1d weight (out, in, 1) -> (out, in) view"""
    return p.detach().view(p.shape[0], p.shape[1])


class _Grads:
    """One flat fp32 gradient buffer with a view per parameter (the allreduce payload)."""

    def __init__(self, net):
        params = list(net.named_parameters())
        dev = params[0][1].device
        self.flat = torch.zeros(sum(p.numel() for _, p in params), device=dev)
        self.views, off = {}, 0
        for name, p in params:
            self.views[name] = self.flat[off:off + p.numel()].view(p.shape)
            off += p.numel()

    def w(self, name):
        v = self.views[name]
        return v.view(v.shape[0], v.shape[1]) if v.dim() == 3 else v


class _PlannedTrunk:
    """8 x (1x1 conv + ReLU) with the skip concat after layer 4 (tpose_nerf_network.py:68-72, 256-260): forward keeping every
    activation, backward producing weight / bias / latent gradients and (optionally) the gradient of the PE input.
    Everything is static: activations live in buffers sized for `m_max` rows, and every product of the forward / backward pass
    is a prebuilt `aninerf_gemm` descriptor -- per call only the row count (M, or K and the split factor of a weight gradient) is
    patched and the library entry invoked.  ~35 launches per pass cost ~3 us of host time each instead of ~16 us (descriptor
    construction, operand views, allocations), which is what bounds a 1024-ray training iteration.  The latent code of the field
    (lat_cols = 128) is folded into the bias of layers 0 and 5 through an M = 1 product, so its gradient comes out of the same
    kernel."""

    def __init__(self, sd, prefix, lat_cols, grads, m_max, dev, pe=None, d_pe=None):
        self.W = [_w2(sd[f'{prefix}.{i}.weight']) for i in range(8)]
        self.b = [sd[f'{prefix}.{i}.bias'].detach() for i in range(8)]
        self.gW = [grads.w(f'{prefix}.{i}.weight') for i in range(8)]
        self.gb = [grads.views[f'{prefix}.{i}.bias'] for i in range(8)]
        self.lat_cols, self.hid0, self.m_max, self.dev = lat_cols, 63 + lat_cols, m_max, dev
        z = lambda *sh: torch.zeros(*sh, device=dev)   # noqa: E731
        self.pe = pe if pe is not None else z(m_max, 64)
        self.d_pe = d_pe
        self.H = [None] + [z(m_max, 256) for _ in range(8)]
        self.dz = [z(m_max, 256), z(m_max, 256)]          # dz[0] = the incoming gradient of layer 7's pre-activation
        self.lat, self.g_lat = z(1, 128), z(1, 128)
        self.beff = {0: z(1, 256), 5: z(1, 256)}
        self.db = z(1, 256)
        self._keep = []
        self.fwd = self._plan_forward()
        self.bwd = {}

    # -- descriptors -----------------------------------------------------------------------------------------------------
    def _gemm(self, segs, out, bias=None, relu=False, relu_mask=None, accumulate=False, dyn=0):
        g = _lib.Gemm()
        g.n_seg = len(segs)
        for i, (a, b) in enumerate(segs):
            sg = g.seg[i]
            sg.A, sg.a_row_stride, sg.a_k_stride = a.ptr, a.rs, a.ks
            sg.B, sg.b_row_stride, sg.b_k_stride = b.ptr, b.rs, b.ks
            sg.K = a.k
        g.M, g.N, g.C, g.ldc = out.shape[0], out.shape[1], out.data_ptr(), out.stride(0)
        if bias is not None:
            g.bias = bias.data_ptr()
        if relu_mask is not None:
            g.relu_mask, g.ld_mask = relu_mask.data_ptr(), relu_mask.stride(0)
        g.relu, g.accumulate, g.split_k = int(relu), int(accumulate), 1
        self._keep.append((segs, out, bias, relu_mask))
        return (0, g, C.byref(g), dyn)            # dyn: 0 static, 1 patch M, 2 patch K of segment 0 (+ split-K)

    def _colsum(self, x, out, accumulate):
        self._keep.append((x, out))
        return (1, (x.data_ptr(), x.stride(0), x.shape[1], out.data_ptr(), int(accumulate)), None, 1)

    def _plan_forward(self):
        ops, H, pe = [], self.H, self.pe
        for l in range(8):
            bias = self.b[l]
            if self.lat_cols and l in (0, 5):
                bias = self.beff[l]
                ops.append(self._gemm([(Op(self.lat), Op(self.W[l][:, 63:63 + self.lat_cols]))], bias, bias=self.b[l]))
            if l == 0:
                segs = [(Op(pe[:, :63]), Op(self.W[0][:, :63]))]
            elif l == 5:
                segs = [(Op(pe[:, :63]), Op(self.W[5][:, :63])), (Op(H[5]), Op(self.W[5][:, self.hid0:]))]
            else:
                segs = [(Op(H[l]), Op(self.W[l]))]
            ops.append(self._gemm(segs, H[l + 1], bias=bias, relu=True, dyn=1))
        return ops

    def _plan_backward(self, want_dpe, wgrad):
        ops, H, pe, dz = [], self.H, self.pe, self.dz
        cur = 0
        for l in range(7, -1, -1):
            W, gW, dZ = self.W[l], self.gW[l], dz[cur]
            if wgrad:
                if l in (0, 5):
                    ops.append(self._gemm([(Op(dZ).T, Op(pe[:, :63]).T)], gW[:, :63], accumulate=True, dyn=2))
                if l == 5:
                    ops.append(self._gemm([(Op(dZ).T, Op(H[5]).T)], gW[:, self.hid0:], accumulate=True, dyn=2))
                elif l != 0:
                    ops.append(self._gemm([(Op(dZ).T, Op(H[l]).T)], gW, accumulate=True, dyn=2))
                if self.lat_cols and l in (0, 5):
                    Wl = W[:, 63:63 + self.lat_cols]
                    ops.append(self._colsum(dZ, self.db, False))
                    ops.append(self._colsum(dZ, self.gb[l], True))
                    ops.append(self._gemm([(Op(self.db).T, Op(self.lat).T)], gW[:, 63:63 + self.lat_cols], accumulate=True))
                    ops.append(self._gemm([(Op(self.db), Op(Wl).T)], self.g_lat, accumulate=(l == 0)))     # layer 5 comes first
                else:
                    ops.append(self._colsum(dZ, self.gb[l], True))
            if want_dpe and l in (0, 5):
                ops.append(self._gemm([(Op(dZ), Op(W[:, :63]).T)], self.d_pe[:, :63], accumulate=(l == 0), dyn=1))
            if l > 0:
                Wh = W[:, self.hid0:] if l == 5 else W
                ops.append(self._gemm([(Op(dZ), Op(Wh).T)], dz[1 - cur], relu_mask=H[l], dyn=1))
                cur = 1 - cur
        return ops

    # -- execution -------------------------------------------------------------------------------------------------------
    def _run(self, ops, m):
        L = _lib.lib()
        gemm, colsum = L.aninerf_gemm_x3, L.aninerf_colsum
        st = T._st(self.dev)
        sk = T.split_for(m)
        ws = T._ws.get(max(sk * 256 * 447 * 4, ((m + 255) // 256) * 256 * 4), self.dev)
        wsp = _lib.ptr(ws)
        for kind, a, ref, dyn in ops:
            if kind == 0:
                nbytes = 0
                if dyn == 1:
                    a.M = m
                elif dyn == 2:
                    a.seg[0].K = m
                    a.split_k = sk
                    nbytes = sk * a.M * a.N * 4 if sk > 1 else 0
                rc = gemm(ref, wsp, nbytes, st)
            else:
                x, ld, n_cols, out, acc = a
                rc = colsum(x, ld, m, n_cols, out, acc, wsp, ((m + 255) // 256) * n_cols * 4, st)
            if rc:
                _lib.check(rc)

    def forward(self, m, lat):
        """The PE of the m rows is in self.pe[:m]; returns H8 (view of the static buffer)."""
        if self.lat_cols:
            self.lat.copy_(lat)
        self._run(self.fwd, m)
        return self.H[8][:m]

    def backward(self, m, g_lat_row, want_dpe, wgrad=True):
        """self.dz[0][:m] holds the gradient of layer 7's pre-activation; the PE gradient lands in self.d_pe[:m]."""
        key = (bool(want_dpe), bool(wgrad))
        if key not in self.bwd:
            self.bwd[key] = self._plan_backward(*key)
        self._run(self.bwd[key], m)
        if self.lat_cols and wgrad and g_lat_row is not None:
            g_lat_row.add_(self.g_lat)


class TrainStep:
    """Forward + backward of one training batch on the library's kernels; holds no state between calls except scratch."""

    def __init__(self, net, cfg=None):
        self.net = net
        self.cfg = cfg if cfg is not None else getattr(net, 'cfg', None) or config.global_cfg()
        self._renderer = Renderer(net, self.cfg)
        self._ws = None
        self._plan = None
        self._generation = 0          # bumped by every forward pass: the static activation buffers belong to the newest one

    def _ensure_plan(self, m, dev, sd):
        """Static state of the step: the flat gradient buffer and the three planned trunks (posed / canonical blend-weight
        field, NeRF trunk), sized for the ACTIVE row count m of the batch (rounded up; it only grows) -- not for the nominal
        n_rays * N_samples rows, of which a training batch keeps a few per cent.  Rebuilt when the capacity, device or
        parameter storage changes."""
        ptrs = tuple(p.data_ptr() for p in sd.values())
        if self._plan is not None and self._plan['ptrs'] == ptrs and self._plan['dev'] == str(dev) and self._plan['cap'] >= m:
            return self._plan
        cap = max(4096, (m + 4095) // 4096 * 4096, self._plan['cap'] if self._plan is not None else 0)
        self._plan = None                                    # release the old buffers first
        G = _Grads(self.net)
        d_pe = torch.zeros(cap, 64, device=dev)
        tp = _PlannedTrunk(sd, 'bw_linears', 128, G, cap, dev)
        tc = _PlannedTrunk(sd, 'bw_linears', 128, G, cap, dev, d_pe=d_pe)
        tn = _PlannedTrunk(sd, 'tpose_human.pts_linears', 0, G, cap, dev, pe=tc.pe, d_pe=d_pe)
        self._plan = {'ptrs': ptrs, 'dev': str(dev), 'cap': cap, 'G': G, 'd_pe': d_pe, 'trunk_p': tp, 'trunk_c': tc, 'trunk_n': tn}
        return self._plan

    def _workspace(self, nbytes, dev):
        if self._ws is None or self._ws.numel() < nbytes or self._ws.device != dev:
            self._ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        return self._ws

    @torch.no_grad()
    def run(self, batch, t_rand=None):
        """Returns (ret, stats, grads): ret = the Renderer.render training contract (device tensors), stats = device scalars
        bw_loss / img_loss / loss, grads = _Grads (d loss / d parameter)."""
        s = self.forward_pass(batch, t_rand)
        losses, d_rgb_map, d_pbw, d_tbw = self.loss_gradients(s, batch)
        G = self.backward_pass(s, d_rgb_map, d_pbw, d_tbw)
        stats = {'bw_loss': losses[0], 'img_loss': losses[1], 'loss': losses[0] + losses[1]}
        return self.contract(s), stats, G

    @staticmethod
    def contract(s):
        """The Renderer.render training contract (tpose_renderer.py:159-186) of a forward pass, device tensors."""
        R, n = s['R'], s['n']
        selb = s['sel'].bool()
        return {'rgb_map': s['rgb_map'].view(1, R, 3), 'acc_map': s['acc_map'].view(1, R), 'depth_map': s['depth_map'].view(1, R),
                'raw': s['raw'].view(1, n, 4), 'pbw': s['pbw'][selb].view(1, -1, 24), 'tbw': s['tbw'][selb].view(1, -1, 24)}

    @torch.no_grad()
    def forward_pass(self, batch, t_rand=None):
        """Network.forward + raw2outputs of one training batch, keeping every activation the backward pass needs (in the
        step's static buffers: a later forward pass invalidates this one's state, `backward_pass` checks)."""
        cfg, net, L = self.cfg, self.net, _lib.lib()
        sd = dict(net.named_parameters())
        if 'novel_pose_bw.bw_fc.weight' in sd and config.get(cfg, 'test_novel_pose'):
            raise _lib.AninerfError('tpose_trainer trains the frame-indexed fields; the novel-pose stage is aninerf_animation_trainer')
        ray_o = batch['ray_o']
        _lib.require_cuda(ray_o, "batch['ray_o']")
        dev = ray_o.device
        st = _lib.stream_ptr(dev)
        T.begin_step(dev)
        R = ray_o.shape[1]
        S = int(config.get(cfg, 'N_samples'))
        n = R * S
        o, d = _lib.f32c(ray_o.reshape(-1, 3)), _lib.f32c(batch['ray_d'].reshape(-1, 3))
        near, far = _lib.f32c(batch['near'].reshape(-1)), _lib.f32c(batch['far'].reshape(-1))
        fr, keep = _frame_struct(batch, need_tbw=True)
        pr = _lib.RenderParams(n_samples=S, chunk_rays=_lib.CHUNK_RAYS, norm_th=float(config.get(cfg, 'norm_th')),
                               white_bkgd=int(bool(config.get(cfg, 'white_bkgd'))), novel_pose=0, want_bw=1, bw_precision=3, nerf_precision=3)
        tv = self._renderer._tv(S, dev)
        tr = _lib.f32c(t_rand.reshape(R, S).to(dev)) if t_rand is not None else None
        latent_index = int(torch.as_tensor(batch['latent_index']).reshape(-1)[0])

        # ---- front end: samples -> pose space -> pnorm mask -> per-chunk argmin forcing -> compaction -----------------
        n_chunks = (R + _lib.CHUNK_RAYS - 1) // _lib.CHUNK_RAYS
        index = torch.empty(n, dtype=torch.int32, device=dev)
        ppts_all, vd_all, dists_all = torch.empty(n, 3, device=dev), torch.empty(n, 3, device=dev), torch.empty(n, device=dev)
        z_vals = torch.empty(R, S, device=dev)
        n_active = torch.zeros(1, dtype=torch.int32, device=dev)
        chunk_offsets = torch.zeros(n_chunks + 1, dtype=torch.int32, device=dev)
        pv = int(fr.pbw_dims[0]) * fr.pbw_dims[1] * fr.pbw_dims[2]
        ws_bytes = L.aninerf_front_end_workspace_bytes(R, S, pv)
        ws = self._workspace(ws_bytes, dev)
        _lib.check(L.aninerf_front_end(C.byref(fr), C.byref(pr), _lib.ptr(o), _lib.ptr(d), _lib.ptr(near), _lib.ptr(far), _lib.ptr(tv), _lib.ptr(tr),
                                       R, _lib.ptr(index), _lib.ptr(ppts_all), _lib.ptr(vd_all), _lib.ptr(dists_all), _lib.ptr(z_vals),
                                       _lib.ptr(n_active), _lib.ptr(chunk_offsets), _lib.ptr(ws), ws_bytes, st))
        m = int(n_active.item())                      # the reference syncs here too (boolean indexing, tpose_nerf_network.py:155-157)
        index, ppts, viewdir, dists = index[:m], ppts_all[:m], vd_all[:m], dists_all[:m]

        plan = self._ensure_plan(m, dev, sd)
        self._generation += 1
        trunk_p, trunk_c, trunk_n = plan['trunk_p'], plan['trunk_c'], plan['trunk_n']
        A = _lib.f32c(batch['A'].reshape(24, 4, 4))
        pvol, tvol = _lib.f32c(batch['pbw'][0]), _lib.f32c(batch['tbw'][0])
        pb, tb = _lib.f32c(batch['pbounds'].reshape(2, 3)), _lib.f32c(batch['tbounds'].reshape(2, 3))
        bw_lat, nf_lat = sd['bw_latent.weight'].detach(), sd['tpose_human.nf_latent.weight'].detach()
        lat_p, lat_c, lat_n = bw_lat[latent_index + 1:latent_index + 2], bw_lat[0:1], nf_lat[latent_index:latent_index + 1]
        Wfc, bfc = _w2(sd['bw_fc.weight']), sd['bw_fc.bias'].detach()

        def e(*shape):
            return torch.empty(*shape, device=dev)

        # ---- blend-weight field at the posed points + inverse LBS (tpose_nerf_network.py:79-100) ------------------------
        T.pe_forward(ppts, 10, trunk_p.pe[:m])
        h8p = trunk_p.forward(m, lat_p)
        delta_p = T.gemm([(Op(h8p), Op(Wfc))], e(m, 24), bias=bfc)
        init_p = T.sample_volume(ppts, pvol, pb, e(m, 25))
        pbw = T.bw_softmax_forward(init_p, delta_p, e(m, 24))
        tpts = T.inverse_lbs(ppts, pbw, A, e(m, 3))
        # ---- blend-weight field at the canonical points, latent index 0 (:163-170) --------------------------------------
        T.pe_forward(tpts, 10, trunk_c.pe[:m])
        h8c = trunk_c.forward(m, lat_c)
        delta_c = T.gemm([(Op(h8c), Op(Wfc))], e(m, 24), bias=bfc)
        init_t = T.sample_volume(tpts, tvol, tb, e(m, 25))
        tbw = T.bw_softmax_forward(init_t, delta_c, e(m, 24))
        # ---- canonical NeRF field (:252-275) ----------------------------------------------------------------------------
        p = 'tpose_human.'
        h8n = trunk_n.forward(m, None)          # shares the PE(tpose) buffer of the canonical blend-weight trunk
        Wa, Wf, Wl, Wv, Wr = (_w2(sd[p + k + '.weight']) for k in ('alpha_fc', 'feature_fc', 'latent_fc', 'view_fc', 'rgb_fc'))
        ba, bf, bl, bv, br = (sd[p + k + '.bias'].detach() for k in ('alpha_fc', 'feature_fc', 'latent_fc', 'view_fc', 'rgb_fc'))
        sigma = T.gemm([(Op(h8n), Op(Wa))], e(m, 1), bias=ba)
        feat = T.gemm([(Op(h8n), Op(Wf))], e(m, 256), bias=bf)
        bl_eff = T.gemm([(Op(lat_n), Op(Wl[:, 256:]))], e(1, 256), bias=bl)
        f2 = T.gemm([(Op(feat), Op(Wl[:, :256]))], e(m, 256), bias=bl_eff)
        pe_v = T.pe_forward(viewdir, 4, e(m, 32))
        hv = T.gemm([(Op(f2), Op(Wv[:, :256])), (Op(pe_v[:, :27]), Op(Wv[:, 256:]))], e(m, 128), bias=bv, relu=True)
        rgb = T.gemm([(Op(hv), Op(Wr))], e(m, 3), bias=br)
        # ---- tail, compositing, losses ------------------------------------------------------------------------------------
        raw = torch.zeros(n, 4, device=dev)
        sigma_masked = e(m)
        _lib.check(L.aninerf_nerf_tail_forward(_lib.ptr(sigma), _lib.ptr(rgb), _lib.ptr(tpts), _lib.ptr(tb), _lib.ptr(dists), _lib.ptr(index), m,
                                               _lib.ptr(raw), _lib.ptr(sigma_masked), st))
        rgb_map, acc_map, depth_map = e(R, 3), e(R), e(R)
        _lib.check(L.aninerf_composite(_lib.ptr(raw), _lib.ptr(z_vals), R, S, pr.white_bkgd, _lib.ptr(rgb_map), _lib.ptr(acc_map),
                                       _lib.ptr(depth_map), None, None, st))
        sel = torch.empty(m, dtype=torch.uint8, device=dev)
        n_sel = torch.zeros(1, dtype=torch.int32, device=dev)
        _lib.check(L.aninerf_select_rows(_lib.ptr(sigma_masked), _lib.ptr(chunk_offsets), n_chunks, float(config.get(cfg, 'train_th')),
                                         _lib.ptr(sel), _lib.ptr(n_sel), st))
        self._keep = (keep, o, d, near, far, tr, A, pvol, tvol, pb, tb)
        T.end_step(dev)
        return dict(gen=self._generation, R=R, S=S, n=n, m=m, dev=dev, white=pr.white_bkgd, latent_index=latent_index, sd=sd, plan=plan,
                    index=index, ppts=ppts, viewdir=viewdir, dists=dists, z_vals=z_vals, chunk_offsets=chunk_offsets, A=A, pvol=pvol, tvol=tvol,
                    pb=pb, tb=tb, lat_n=lat_n, h8p=h8p, h8c=h8c, h8n=h8n, init_p=init_p, init_t=init_t, pbw=pbw, tbw=tbw, tpts=tpts,
                    feat=feat, f2=f2, pe_v=pe_v, hv=hv, raw=raw, sigma_masked=sigma_masked, rgb_map=rgb_map, acc_map=acc_map,
                    depth_map=depth_map, sel=sel, n_sel=n_sel)

    @torch.no_grad()
    def loss_gradients(self, s, batch):
        """The two losses of tpose_trainer.py:48-63 and their gradients w.r.t. rgb_map (R,3), pbw and tbw (m,24; zero on the
        rows `alpha_ind` drops).  Returns (losses (2,) = [bw_loss, img_loss], d_rgb_map, d_pbw, d_tbw)."""
        L, dev, m, R = _lib.lib(), s['dev'], s['m'], s['R']
        st = _lib.stream_ptr(dev)
        pbw, tbw, sel, n_sel, rgb_map = s['pbw'], s['tbw'], s['sel'], s['n_sel'], s['rgb_map']

        def e(*shape):
            return torch.empty(*shape, device=dev)
        losses = torch.zeros(2, device=dev)
        d_pbw, d_tbw = e(m, 24), e(m, 24)
        _lib.check(L.aninerf_bw_loss(_lib.ptr(pbw), _lib.ptr(tbw), _lib.ptr(sel), _lib.ptr(n_sel), m, _lib.ptr(losses[0:1]), _lib.ptr(d_pbw),
                                     _lib.ptr(d_tbw), st))
        rgb_gt = _lib.f32c(batch['rgb'].reshape(R, 3))
        mask = batch['mask_at_box'].reshape(R).to(torch.uint8).contiguous()
        d_rgb_map = e(R, 3)
        _lib.check(L.aninerf_img_loss(_lib.ptr(rgb_map), _lib.ptr(rgb_gt), _lib.ptr(mask), R, _lib.ptr(losses[1:2]), _lib.ptr(d_rgb_map), st))

        self._keep_loss = (rgb_gt, mask)
        return losses, d_rgb_map, d_pbw, d_tbw

    @torch.no_grad()
    def backward_pass(self, s, d_rgb_map, d_pbw, d_tbw):
        """d loss / d parameter for gradients arriving on rgb_map (R,3), pbw (m,24) and tbw (m,24) of forward pass `s`
        (d_pbw is updated in place).  Returns the step's _Grads (its flat buffer is overwritten by the next backward pass)."""
        if s['gen'] != self._generation:
            raise _lib.AninerfError('backward of a stale forward pass: the training step keeps ONE set of activation buffers, so a '
                                    'forward pass must be followed by its backward before the next forward (no gradient accumulation '
                                    'over several forwards)')
        cfg, L = self.cfg, _lib.lib()
        dev, m, n, R, S, sd, plan = s['dev'], s['m'], s['n'], s['R'], s['S'], s['sd'], s['plan']
        st = _lib.stream_ptr(dev)
        T.begin_step(dev)
        G = plan['G']
        G.flat.zero_()
        trunk_p, trunk_c, trunk_n, d_pe_all = plan['trunk_p'], plan['trunk_c'], plan['trunk_n'], plan['d_pe']
        latent_index, A, tvol, tb, lat_n = s['latent_index'], s['A'], s['tvol'], s['tb'], s['lat_n']
        index, dists, tpts, raw, sigma_masked = s['index'], s['dists'], s['tpts'], s['raw'], s['sigma_masked']
        h8p, h8c, h8n, init_p, init_t, pbw, tbw = s['h8p'], s['h8c'], s['h8n'], s['init_p'], s['init_t'], s['pbw'], s['tbw']
        feat, f2, pe_v, hv = s['feat'], s['f2'], s['pe_v'], s['hv']
        p = 'tpose_human.'
        Wfc = _w2(sd['bw_fc.weight'])
        Wa, Wf, Wl, Wv, Wr = (_w2(sd[p + k + '.weight']) for k in ('alpha_fc', 'feature_fc', 'latent_fc', 'view_fc', 'rgb_fc'))
        white = s['white']

        def e(*shape):
            return torch.empty(*shape, device=dev)

        # ================================= backward =========================================================================
        d_raw = e(n, 4)
        _lib.check(L.aninerf_composite_backward(_lib.ptr(raw), _lib.ptr(d_rgb_map), R, S, white, _lib.ptr(d_raw), st))
        d_sigma, d_rgb = e(m, 1), e(m, 3)
        _lib.check(L.aninerf_nerf_tail_backward(_lib.ptr(d_raw), _lib.ptr(raw), _lib.ptr(index), _lib.ptr(sigma_masked), _lib.ptr(tpts), _lib.ptr(tb),
                                                _lib.ptr(dists), m, _lib.ptr(d_sigma), _lib.ptr(d_rgb), st))
        sk = T.split_for(m)
        gw, gv = G.w, G.views
        # NeRF heads
        T.gemm([(Op(d_rgb).T, Op(hv).T)], gw(p + 'rgb_fc.weight'), accumulate=True, split_k=sk)
        T.colsum(d_rgb, gv[p + 'rgb_fc.bias'], accumulate=True)
        dzv = T.gemm([(Op(d_rgb), Op(Wr).T)], e(m, 128), relu_mask=hv)
        gWv = gw(p + 'view_fc.weight')
        T.gemm([(Op(dzv).T, Op(f2).T)], gWv[:, :256], accumulate=True, split_k=sk)
        T.gemm([(Op(dzv).T, Op(pe_v[:, :27]).T)], gWv[:, 256:], accumulate=True, split_k=sk)
        T.colsum(dzv, gv[p + 'view_fc.bias'], accumulate=True)
        d_f2 = T.gemm([(Op(dzv), Op(Wv[:, :256]).T)], e(m, 256))
        gWl = gw(p + 'latent_fc.weight')
        T.gemm([(Op(d_f2).T, Op(feat).T)], gWl[:, :256], accumulate=True, split_k=sk)
        dbl = T.colsum(d_f2, e(1, 256))
        gv[p + 'latent_fc.bias'].add_(dbl.view(-1))
        T.gemm([(Op(dbl).T, Op(lat_n).T)], gWl[:, 256:], accumulate=True)
        T.gemm([(Op(dbl), Op(Wl[:, 256:]).T)], gv[p + 'nf_latent.weight'][latent_index:latent_index + 1], accumulate=True)
        d_feat = T.gemm([(Op(d_f2), Op(Wl[:, :256]).T)], e(m, 256))
        T.gemm([(Op(d_feat).T, Op(h8n).T)], gw(p + 'feature_fc.weight'), accumulate=True, split_k=sk)
        T.colsum(d_feat, gv[p + 'feature_fc.bias'], accumulate=True)
        T.gemm([(Op(d_sigma).T, Op(h8n).T)], gw(p + 'alpha_fc.weight'), accumulate=True, split_k=sk)
        T.colsum(d_sigma, gv[p + 'alpha_fc.bias'], accumulate=True)
        dz = T.gemm([(Op(d_feat), Op(Wf).T)], trunk_n.dz[0][:m])
        T.gemm([(Op(d_sigma), Op(Wa).T)], dz, accumulate=True, relu_mask=h8n)
        d_pe = d_pe_all[:m]
        trunk_n.backward(m, None, True)
        d_tpts = T.pe_backward(tpts, d_pe, 10, e(m, 3), False)
        # canonical blend-weight field
        d_delta, d_init = e(m, 24), e(m, 24)
        T.bw_softmax_backward(init_t, tbw, d_tbw, d_delta, d_init)
        gWfc, gbfc = gw('bw_fc.weight'), gv['bw_fc.bias']
        T.gemm([(Op(d_delta).T, Op(h8c).T)], gWfc, accumulate=True, split_k=sk)
        T.colsum(d_delta, gbfc, accumulate=True)
        T.gemm([(Op(d_delta), Op(Wfc).T)], trunk_c.dz[0][:m], relu_mask=h8c)
        g_bw_lat = gv['bw_latent.weight']
        trunk_c.backward(m, g_bw_lat[0:1], True)
        T.pe_backward(tpts, d_pe, 10, d_tpts, True)
        T.sample_volume_backward(tpts, tvol, tb, d_init, d_tpts, True)
        # inverse LBS: d tpose -> d pbw (added to the bw-loss gradient)
        T.inverse_lbs_backward(pbw, A, tpts, d_tpts, d_pbw, True)
        # posed blend-weight field (same parameters: gradients accumulate)
        T.bw_softmax_backward(init_p, pbw, d_pbw, d_delta, None)
        T.gemm([(Op(d_delta).T, Op(h8p).T)], gWfc, accumulate=True, split_k=sk)
        T.colsum(d_delta, gbfc, accumulate=True)
        T.gemm([(Op(d_delta), Op(Wfc).T)], trunk_p.dz[0][:m], relu_mask=h8p)
        trunk_p.backward(m, g_bw_lat[latent_index + 1:latent_index + 2], False)

        T.end_step(dev)
        return G

def _param_views(flat, shapes):
    out, off = [], 0
    for sh in shapes:
        k = 1
        for v in sh:
            k *= v
        out.append(flat[off:off + k].view(sh))
        off += k
    return out


class _LossFn(torch.autograd.Function):
    """Carries the kernel-computed gradients into torch.autograd: loss.backward() adds g * dloss/dparam to param.grad.
    The step's flat gradient buffer is scratch that the next step overwrites, so forward keeps its own copy (5 MB): a second
    forward before this loss's backward (validation between forward and backward, several losses alive) cannot corrupt it."""

    @staticmethod
    def forward(ctx, loss, flat_grad, *params):
        ctx.flat = flat_grad.clone()
        ctx.shapes = [p.shape for p in params]
        return loss.clone()

    @staticmethod
    def backward(ctx, g):
        scaled = ctx.flat * g                      # ONE kernel for all 46 tensors; the per-parameter gradients are views of it
        return (None, None, *_param_views(scaled, ctx.shapes))


class _RenderFn(torch.autograd.Function):
    """`Renderer.render` in training mode as ONE autograd node: forward = TrainStep.forward_pass, backward =
    TrainStep.backward_pass for gradients arriving on rgb_map, pbw and tbw -- the three outputs the reference's losses read
    (tpose_trainer.py:48-63).  acc_map / depth_map / raw are returned without a graph (no reference loss uses them)."""

    @staticmethod
    def forward(ctx, step, batch, t_rand, *params):
        s = step.forward_pass(batch, t_rand)
        ret = TrainStep.contract(s)
        ctx.step, ctx.state = step, s
        ctx.shapes = [p.shape for p in params]
        ctx.mark_non_differentiable(ret['acc_map'], ret['depth_map'], ret['raw'])
        return ret['rgb_map'], ret['acc_map'], ret['depth_map'], ret['raw'], ret['pbw'], ret['tbw']

    @staticmethod
    def backward(ctx, g_rgb, g_acc, g_depth, g_raw, g_pbw, g_tbw):
        s = ctx.state
        dev, m, R = s['dev'], s['m'], s['R']
        selb = s['sel'].bool()
        d_rgb = _lib.f32c(g_rgb.reshape(R, 3)) if g_rgb is not None else torch.zeros(R, 3, device=dev)
        d_pbw, d_tbw = torch.zeros(m, 24, device=dev), torch.zeros(m, 24, device=dev)
        if g_pbw is not None:
            d_pbw[selb] = g_pbw.reshape(-1, 24)
        if g_tbw is not None:
            d_tbw[selb] = g_tbw.reshape(-1, 24)
        G = ctx.step.backward_pass(s, d_rgb, d_pbw, d_tbw)
        return (None, None, None, *_param_views(G.flat.clone(), ctx.shapes))


def render_with_grad(renderer, batch, t_rand=None):
    """tpose_renderer.Renderer.render when a gradient is required (tpose_renderer.py:154-155 keeps the device tensors and their
    graph): the reference's own `NetworkWrapper` (lib/train/trainers/tpose_trainer.py:21-73) can sit on the drop-in renderer."""
    cfg, net = renderer.cfg, renderer.net
    step = renderer.__dict__.get('_train_step')
    if step is None:
        step = renderer.__dict__['_train_step'] = TrainStep(net, cfg)
    if t_rand is None and config.get(cfg, 'perturb') > 0. and net.training:
        t_rand = torch.rand(1, batch['ray_o'].shape[1], int(config.get(cfg, 'N_samples')))     # CPU generator, tpose_renderer.py:35
    params = [p for _, p in net.named_parameters()]
    rgb_map, acc_map, depth_map, raw, pbw, tbw = _RenderFn.apply(step, batch, t_rand, *params)
    return {'rgb_map': rgb_map, 'acc_map': acc_map, 'depth_map': depth_map, 'raw': raw, 'pbw': pbw, 'tbw': tbw}


class NetworkWrapper(nn.Module):
    """tpose_trainer.NetworkWrapper: forward(batch) -> (ret, loss, scalar_stats, image_stats)."""

    def __init__(self, net, cfg=None):
        super().__init__()
        self.net = net
        self.cfg = cfg if cfg is not None else getattr(net, 'cfg', None) or config.global_cfg()
        self.renderer = Renderer(net, self.cfg)
        self.__dict__['_step'] = TrainStep(net, self.cfg)

    def forward(self, batch, t_rand=None):
        cfg = self.cfg
        if not (torch.is_grad_enabled() and self.net.training):
            return self._forward_eval(batch)
        if t_rand is None and config.get(cfg, 'perturb') > 0.:
            R = batch['ray_o'].shape[1]
            t_rand = torch.rand(1, R, int(config.get(cfg, 'N_samples')))       # CPU generator, as tpose_renderer.py:35
        ret, stats, G = self.__dict__['_step'].run(batch, t_rand)
        params = [p for _, p in self.net.named_parameters()]
        loss = _LossFn.apply(stats['loss'], G.flat, *params)
        scalar_stats = {'bw_loss': stats['bw_loss'], 'img_loss': stats['img_loss'], 'loss': loss}
        return ret, loss, scalar_stats, {}

    @torch.no_grad()
    def _forward_eval(self, batch):
        """Validation (Trainer.val, lib/train/trainers/trainer.py:84-115: eval mode, no_grad, a whole mask_at_box image of ~1e5
        rays): the fused render path -- chunk-safe, no activation plan, no backward kernels -- plus the two losses."""
        import torch.nn.functional as F
        r = self.renderer
        out = r.render_device(batch, want_bw=True)
        sel, n_sel = r.select_rows(out)
        k = int(n_sel.item())
        pbw, tbw = r.gather_selected(out, sel, k)
        R = batch['ray_o'].shape[1]
        S = int(config.get(self.cfg, 'N_samples'))
        ret = {'rgb_map': out['rgb_map'].view(1, R, 3), 'acc_map': out['acc_map'].view(1, R), 'depth_map': out['depth_map'].view(1, R),
               'raw': out['raw'].view(1, R * S, 4), 'pbw': pbw.view(1, k, 24), 'tbw': tbw.view(1, k, 24)}
        bw_loss = F.smooth_l1_loss(ret['pbw'], ret['tbw'])
        mask = batch['mask_at_box'].reshape(1, R).bool()
        img_loss = torch.mean((ret['rgb_map'][mask] - batch['rgb'].reshape(1, R, 3)[mask]) ** 2)
        loss = bw_loss + img_loss
        return ret, loss, {'bw_loss': bw_loss, 'img_loss': img_loss, 'loss': loss}, {}


def allreduce_gradients(net, world_size: int, group=None):
    """Data-parallel gradient average (what DistributedDataParallel does in trainer.py:13-19) as ONE collective over a flat
    buffer: NCCL all-reduce(sum) of 1 274 652 floats, then 1/world."""
    if world_size <= 1:
        return
    import torch.distributed as dist
    params = [p for p in net.parameters() if p.grad is not None]
    if not params:
        return
    flat = torch.cat([p.grad.reshape(-1) for p in params])
    dist.all_reduce(flat, group=group)
    flat.mul_(1.0 / world_size)
    off = 0
    for p in params:
        k = p.numel()
        p.grad.copy_(flat[off:off + k].view_as(p.grad))
        off += k


def train_iteration(wrapper: NetworkWrapper, batch, optimizer, world_size: int = 1, t_rand=None, clip: float = 40.0, group=None):
    """One iteration of Trainer.train (trainer.py:62-66) -- zero_grad, forward, backward, [DDP mean], clip_grad_value_(40), step --
    without the round trip through torch.autograd: the step's flat gradient buffer IS the all-reduce payload (one NCCL all-reduce
    of 1 274 652 floats), is clipped with one kernel, and its per-parameter views become `.grad`.  Same numbers as
    `loss.backward()` on `NetworkWrapper.forward` (tests/test_gpu_train.py), ~1 ms less host time per iteration.
    NOTE: after this call every `p.grad` is a VIEW of the step's flat gradient buffer -- scratch memory that the next
    `train_iteration` / `NetworkWrapper.forward` zeroes and overwrites.  Read or copy gradients before the next step."""
    cfg = wrapper.cfg
    if t_rand is None and config.get(cfg, 'perturb') > 0. and wrapper.net.training:
        t_rand = torch.rand(1, batch['ray_o'].shape[1], int(config.get(cfg, 'N_samples')))
    ret, stats, G = wrapper.__dict__['_step'].run(batch, t_rand)
    if world_size > 1:
        import torch.distributed as dist
        dist.all_reduce(G.flat, group=group)
        G.flat.mul_(1.0 / world_size)
    G.flat.clamp_(-clip, clip)
    for name, p in wrapper.net.named_parameters():
        p.grad = G.views[name]
    optimizer.step()
    return ret, stats
