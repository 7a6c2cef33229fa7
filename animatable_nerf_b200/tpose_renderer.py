"""Drop-in `Renderer` for `lib/networks/renderer/tpose_renderer.py` (select it with
`renderer_module` / `renderer_path`).

`Renderer(net).render(batch)` keeps the reference contract (tpose_renderer.py:159-186): the batch
schema of `lib/datasets/tpose_dataset.py`:236-277 in, a dict with `rgb_map (1,R,3)`, `acc_map (1,R)`,
`depth_map (1,R)`, `raw (1,R*S,4)`, `pbw`/`tbw (1,n'',24)` out, CPU tensors when no gradient is
required.  The whole frame goes through ONE call of the fused C-ABI entry `aninerf_render_rays`;
the reference's 2048-ray chunk loop survives only as the semantic unit of the per-chunk
argmin / argmax forcing (tpose_nerf_network.py:154, :193-194).
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib, config
from .tpose_nerf_network import _frame_struct


class Renderer:
    def __init__(self, net, cfg=None):
        self.net = net
        self.cfg = cfg if cfg is not None else getattr(net, 'cfg', None) or config.global_cfg()
        self._ws = None
        self._t_vals = {}

    # ---------------------------------------------------------------------------------------
    def _workspace(self, nbytes, device):
        if self._ws is None or self._ws.numel() < nbytes or self._ws.device != device:
            self._ws = torch.empty(nbytes, dtype=torch.uint8, device=device)
        return self._ws

    def _tv(self, S, device):
        key = (S, device)
        if key not in self._t_vals:
            # torch.linspace on the CPU, as tpose_renderer.py:26 computes it (not bit-equal to i/(S-1))
            self._t_vals[key] = torch.linspace(0., 1., steps=S).to(device)
        return self._t_vals[key]

    # ---------------------------------------------------------------------------------------
    FRAME_KEYS_RENDER = ('A', 'R', 'Th', 'pbw', 'pbounds', 'tbounds', 'wbounds', 'latent_index', 'bw_latent_index', 'ray_o', 'ray_d', 'near', 'far',
                         'msks', 'Ks', 'RT', 'H', 'W')

    def to_device(self, batch, device=None, non_blocking=True, pool=None):
        """Host batch -> device batch (the `batch[k] = batch[k].cuda()` loop of run.py:63-66), moving only the keys the
        configured mode READS: in render-only mode (`b200_render_only`) the canonical volume `tbw` (11 MB per frame), `occupancy`,
        `rgb`, ... never reach a kernel and stay on the host.  The dozen small tensors of a batch (bone matrices, bounds, pose,
        indices: a few KB together) travel as ONE staged copy instead of a dozen 10-microsecond `.to()` calls; pinned host tensors
        are copied asynchronously on the current stream.
        pool: a dict owned by the caller; the device tensors (and the staging buffers of the small ones) are kept in it and REUSED by
        the next call with the same pool and shapes -- no allocation in a steady-state frame loop (`render_frames` rotates three)."""
        dev = torch.device(device) if device is not None else next(self.net.parameters()).device
        render_only = bool(config.get(self.cfg, 'b200_render_only'))
        out, small = {}, []
        for k, v in batch.items():
            if not torch.is_tensor(v):
                out[k] = v
            elif not render_only or k in self.FRAME_KEYS_RENDER:
                if v.device.type == 'cpu' and v.numel() * v.element_size() <= 16384 and v.numel() > 0:
                    small.append((k, v))
                elif pool is not None and v.device.type == 'cpu':
                    buf = pool.get(k)
                    if buf is None or buf.shape != v.shape or buf.dtype != v.dtype or buf.device != dev:
                        buf = pool[k] = torch.empty(v.shape, dtype=v.dtype, device=dev)
                    buf.copy_(v, non_blocking=non_blocking)
                    out[k] = buf
                else:
                    out[k] = v.to(dev, non_blocking=non_blocking)
        if small:
            offs, total = [], 0
            for _, v in small:
                offs.append(total)
                total += (v.numel() * v.element_size() + 15) // 16 * 16
            store = pool if pool is not None else self.__dict__
            stage, done = store.get('_stage'), store.get('_stage_done')
            if done is not None:
                done.synchronize()                      # the previous staged copy has left the pinned buffer
            if stage is None or stage.numel() < total:
                stage = store['_stage'] = torch.empty(max(total, 65536), dtype=torch.uint8).pin_memory()
            for (k, v), o in zip(small, offs):
                stage[o:o + v.numel() * v.element_size()].copy_(v.contiguous().view(-1).view(torch.uint8))
            if pool is not None:
                d = pool.get('_stage_dev')
                if d is None or d.numel() < total or d.device != dev:
                    d = pool['_stage_dev'] = torch.empty(stage.numel(), dtype=torch.uint8, device=dev)
                d = d[:total]
                d.copy_(stage[:total], non_blocking=non_blocking)
            else:
                d = stage[:total].to(dev, non_blocking=non_blocking)
            if dev.type == 'cuda':
                done = store['_stage_done'] = torch.cuda.Event()
                done.record(torch.cuda.current_stream(dev))
            for (k, v), o in zip(small, offs):
                out[k] = d[o:o + v.numel() * v.element_size()].view(v.dtype).view(v.shape)
        return out

    @torch.no_grad()
    def render_device(self, batch, t_rand=None, want_bw=None, silhouettes=None, peers=None, keep_raw=False):
        """The fused path, results left on the device.  Returns a dict with rgb_map/acc_map/depth_map
        (and raw, n_active, pbw_all/tbw_all/sigma_masked/chunk_offsets when want_bw).
        silhouettes: (_lib.Silhouettes, keep-alive) from tpose_renderer_mmsk -- cull the samples first.
        peers: _lib.PeerGather from ray_tiles.PeerImage -- the compositing kernel also stores the rows into every rank's image."""
        cfg = self.cfg
        ray_o, ray_d = batch['ray_o'], batch['ray_d']
        _lib.require_cuda(ray_o, "batch['ray_o']")
        dev = ray_o.device
        R = ray_o.shape[1]
        S = int(config.get(cfg, 'N_samples'))
        if want_bw is None:
            want_bw = not bool(config.get(cfg, 'b200_render_only'))
        o, d = _lib.f32c(ray_o.reshape(-1, 3)), _lib.f32c(ray_d.reshape(-1, 3))
        near, far = _lib.f32c(batch['near'].reshape(-1)), _lib.f32c(batch['far'].reshape(-1))
        fr, keep = _frame_struct(batch, need_tbw=want_bw)
        pr = _lib.RenderParams(n_samples=S, chunk_rays=_lib.CHUNK_RAYS, norm_th=float(config.get(cfg, 'norm_th')),
                               white_bkgd=int(bool(config.get(cfg, 'white_bkgd'))),
                               novel_pose=int(bool(config.get(cfg, 'test_novel_pose'))), want_bw=int(want_bw),
                               bw_precision=int(config.get(cfg, 'b200_bw_precision')),
                               nerf_precision=int(config.get(cfg, 'b200_nerf_precision')))
        n = R * S
        n_chunks = (R + _lib.CHUNK_RAYS - 1) // _lib.CHUNK_RAYS
        out = {
            'rgb_map': torch.empty(R, 3, device=dev), 'acc_map': torch.empty(R, device=dev), 'depth_map': torch.empty(R, device=dev),
'n_active': torch.zeros(1, dtype=torch.int32, device=dev),
            'chunk_offsets': torch.zeros(n_chunks + 1, dtype=torch.int32, device=dev),
        }
        if want_bw or keep_raw:
            out['raw'] = torch.empty(n, 4, device=dev)     # dense (n,4): a training-contract output; render-only composites the compact rows
        if want_bw:
            out['pbw_all'] = torch.empty(n, 24, device=dev)
            out['tbw_all'] = torch.empty(n, 24, device=dev)
            out['sigma_masked'] = torch.empty(n, device=dev)
            out['active_index'] = torch.empty(n, dtype=torch.int32, device=dev)
        ro = _lib.RenderOutputs()
        for k in ('rgb_map', 'acc_map', 'depth_map', 'raw', 'pbw_all', 'tbw_all', 'sigma_masked', 'active_index', 'n_active',
                  'chunk_offsets'):
            setattr(ro, k, out[k].data_ptr() if k in out else None)
        pv = int(fr.pbw_dims[0]) * fr.pbw_dims[1] * fr.pbw_dims[2]
        tv = int(fr.tbw_dims[0]) * fr.tbw_dims[1] * fr.tbw_dims[2] if want_bw else 0
        ws_bytes = _lib.lib().aninerf_render_workspace_bytes(R, S, int(want_bw), pv, tv)
        ws = self._workspace(ws_bytes, dev)
        tr = _lib.f32c(t_rand.reshape(R, S)) if t_rand is not None else None
        if peers is not None:
            _lib.check(_lib.lib().aninerf_render_rays_tiled(self.net.packed().handle, C.byref(fr), C.byref(pr),
                                                            C.byref(silhouettes[0]) if silhouettes is not None else None, C.byref(peers),
                                                            _lib.ptr(o), _lib.ptr(d), _lib.ptr(near), _lib.ptr(far), _lib.ptr(self._tv(S, dev)),
                                                            _lib.ptr(tr), R, C.byref(ro), _lib.ptr(ws), ws_bytes, _lib.stream_ptr(dev)))
        elif silhouettes is None:
            _lib.check(_lib.lib().aninerf_render_rays(self.net.packed().handle, C.byref(fr), C.byref(pr), _lib.ptr(o), _lib.ptr(d),
                                                      _lib.ptr(near), _lib.ptr(far), _lib.ptr(self._tv(S, dev)), _lib.ptr(tr), R,
                                                      C.byref(ro), _lib.ptr(ws), ws_bytes, _lib.stream_ptr(dev)))
        else:
            _lib.check(_lib.lib().aninerf_render_rays_culled(self.net.packed().handle, C.byref(fr), C.byref(pr), C.byref(silhouettes[0]),
                                                             _lib.ptr(o), _lib.ptr(d), _lib.ptr(near), _lib.ptr(far),
                                                             _lib.ptr(self._tv(S, dev)), _lib.ptr(tr), R, C.byref(ro), _lib.ptr(ws),
                                                             ws_bytes, _lib.stream_ptr(dev)))
        out['_keep'] = (keep, o, d, near, far, tr, silhouettes)
        return out

    # ---------------------------------------------------------------------------------------
    @torch.no_grad()
    def select_rows(self, out):
        """alpha_ind of tpose_nerf_network.py:192-194 for the whole frame on the device (`aninerf_select_rows`: sigma > train_th
        plus the first arg-max row of every 2048-ray chunk).  out: a want_bw result of render_device.
        Returns (sel uint8 (n,), n_sel int32 (1,)) device tensors; only the first n_active entries of sel are meaningful."""
        dev = out['sigma_masked'].device
        n_chunks = out['chunk_offsets'].numel() - 1
        sel = torch.empty(out['sigma_masked'].numel(), dtype=torch.uint8, device=dev)
        n_sel = torch.zeros(1, dtype=torch.int32, device=dev)
        _lib.check(_lib.lib().aninerf_select_rows(_lib.ptr(out['sigma_masked']), _lib.ptr(out['chunk_offsets']), n_chunks,
                                                  float(config.get(self.cfg, 'train_th')), _lib.ptr(sel), _lib.ptr(n_sel), _lib.stream_ptr(dev)))
        return sel, n_sel

    @torch.no_grad()
    def gather_selected(self, out, sel, n_sel_host):
        """pbw[alpha_ind], tbw[alpha_ind] (tpose_nerf_network.py:195-196) -> two (n_sel, 24) device tensors, rows ascending."""
        dev = sel.device
        n_chunks = out['chunk_offsets'].numel() - 1
        pbw = torch.empty(max(n_sel_host, 1), 24, device=dev)
        tbw = torch.empty(max(n_sel_host, 1), 24, device=dev)
        offs = torch.empty(n_chunks + 1, dtype=torch.int32, device=dev)
        _lib.check(_lib.lib().aninerf_gather_selected_rows(_lib.ptr(sel), _lib.ptr(out['chunk_offsets']), n_chunks, _lib.ptr(out['pbw_all']),
                                                           _lib.ptr(out['tbw_all']), _lib.ptr(pbw), _lib.ptr(tbw), _lib.ptr(offs), _lib.stream_ptr(dev)))
        return pbw[:n_sel_host], tbw[:n_sel_host]

    @staticmethod
    def _to_host(tensors, device):
        """One batched device->host transfer into pinned buffers (torch's caching host allocator recycles them), one sync.
        The reference's per-chunk `.detach().cpu()` (tpose_renderer.py:154-155) goes through pageable memory at a fraction of
        the PCIe rate; the results are ordinary CPU tensors either way."""
        host = {}
        for k, v in tensors.items():
            h = torch.empty(v.shape, dtype=v.dtype, pin_memory=True)
            h.copy_(v, non_blocking=True)
            host[k] = h
        torch.cuda.current_stream(device).synchronize()
        return host

    def render(self, batch):
        """Renderer.render (tpose_renderer.py:159-186).  No gradient attached (evaluation runs under torch.no_grad(), run.py:62):
        CPU tensors (:154-155).  Gradient mode with trainable parameters: device tensors that carry the autograd graph to the
        Network parameters (the contract tpose_trainer.NetworkWrapper relies on, lib/train/trainers/tpose_trainer.py:28)."""
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.net.parameters()):      # = `rgb_map.requires_grad` of tpose_renderer.py:154
            from .tpose_trainer import render_with_grad
            return render_with_grad(self, batch)
        with torch.no_grad():
            return self._render_eval(batch)

    def render_frames(self, host_batches, device=None):
        """The evaluation loop of run.py:59-70 (`for batch in data_loader: batch[k] = batch[k].cuda(); renderer.render(batch)`) as
        a generator over HOST batches: yields `render(to_device(batch))` frame by frame, in order, with
          * the NEXT batch's upload in flight on a second stream while the current frame renders (the 18 MB blend-weight volume of
            a frame takes ~0.35 ms of PCIe time -- a seventh of the frame's kernels), and
          * in the render-only evaluation mode, the PREVIOUS frame's maps downloaded on a third stream while the current frame
            renders: frame i is handed out once frame i+1 has been enqueued, so the device never waits for the host loop.
        Results are identical to calling `render(to_device(batch))` per batch."""
        dev = torch.device(device) if device is not None else next(self.net.parameters()).device
        main = torch.cuda.current_stream(dev)
        side, down = self.__dict__.get('_copy_stream'), self.__dict__.get('_down_stream')
        if side is None:
            side = self.__dict__['_copy_stream'] = torch.cuda.Stream(dev)
            down = self.__dict__['_down_stream'] = torch.cuda.Stream(dev)
        cfg = self.cfg
        lagged = bool(config.get(cfg, 'b200_render_only')) and not (torch.is_grad_enabled() and any(p.requires_grad for p in self.net.parameters()))

        pools = self.__dict__.setdefault('_frame_pools', [{}, {}, {}])      # inputs of the frame ahead, the one rendering, the one downloading
        turn = [0]

        def upload(b):
            begun = torch.cuda.Event()
            begun.record(main)
            side.wait_event(begun)                       # (the pool's previous user -- three frames back -- is complete by then)
            with torch.cuda.stream(side):
                d = self.to_device(b, dev, pool=pools[turn[0] % 3])
                turn[0] += 1
                ev = torch.cuda.Event()
                ev.record(side)
            return d, ev

        def finish(p):
            staged, ev, _keep = p
            ev.synchronize()
            # ordinary CPU tensors, as `.cpu()` returns them; the pinned staging buffers rotate (allocating pinned memory per frame
            # costs a cudaHostAlloc -- 10-50 ms -- whenever the caching host allocator has no block whose stream events are done)
            return {k: v.clone() for k, v in staged.items()}

        rings = self.__dict__.setdefault('_host_rings', [{}, {}, {}])

        it = iter(host_batches)
        first = next(it, None)
        cur = upload(first) if first is not None else None
        pending = None
        while cur is not None:
            nb = next(it, None)
            ahead = upload(nb) if nb is not None else None      # enqueued BEFORE this frame's kernels: the copy engine overlaps them
            d, ev = cur
            main.wait_event(ev)
            for v in d.values():
                if torch.is_tensor(v) and v.is_cuda:
                    v.record_stream(main)
            if lagged:
                with torch.no_grad():
                    R = d['ray_o'].shape[1]
                    t_rand = None
                    if config.get(cfg, 'perturb') > 0. and self.net.training:
                        t_rand = torch.rand(1, R, int(config.get(cfg, 'N_samples'))).to(dev)      # CPU generator, as tpose_renderer.py:35
                    out = self.render_device(d, t_rand=t_rand, want_bw=False)
                    ret = {'rgb_map': out['rgb_map'].view(1, R, 3), 'acc_map': out['acc_map'].view(1, R), 'depth_map': out['depth_map'].view(1, R)}
                    rendered = torch.cuda.Event()
                    rendered.record(main)
                    host, ring = {}, rings[turn[0] % 3]
                    with torch.cuda.stream(down):
                        down.wait_event(rendered)
                        for k, v in ret.items():
                            h = ring.get(k)
                            if h is None or h.shape != v.shape or h.dtype != v.dtype:
                                h = ring[k] = torch.empty(v.shape, dtype=v.dtype, pin_memory=True)
                            h.copy_(v, non_blocking=True)
                            v.record_stream(down)
                            host[k] = h
                        done = torch.cuda.Event()
                        done.record(down)
                if pending is not None:
                    yield finish(pending)
                pending = (host, done, (d, out))
            else:
                yield self.render(d)
            cur = ahead
        if pending is not None:
            yield finish(pending)

    def _render_eval(self, batch):
        cfg = self.cfg
        ray_o = batch['ray_o']
        R = ray_o.shape[1]
        dev = ray_o.device
        S = int(config.get(cfg, 'N_samples'))
        t_rand = None
        if config.get(cfg, 'perturb') > 0. and self.net.training:
            # the reference draws the jitter on the CPU generator (tpose_renderer.py:35)
            t_rand = torch.rand(1, R, S).to(dev)
        render_only = bool(config.get(cfg, 'b200_render_only'))
        out = self.render_device(batch, t_rand=t_rand, want_bw=not render_only)
        ret = {'rgb_map': out['rgb_map'].view(1, R, 3), 'acc_map': out['acc_map'].view(1, R), 'depth_map': out['depth_map'].view(1, R)}
        if render_only:
            return self._to_host(ret, dev)
        sel, n_sel = self.select_rows(out)
        ret['raw'] = out['raw'].view(1, R * S, 4)
        ret['_n_sel'] = n_sel
        host = self._to_host(ret, dev)                       # maps + raw + the selected-row count: one sync
        k = int(host.pop('_n_sel')[0])
        pbw, tbw = self.gather_selected(out, sel, k)
        host.update(self._to_host({'pbw': pbw.view(1, k, 24), 'tbw': tbw.view(1, k, 24)}, dev))
        return host
