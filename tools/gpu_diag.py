"""Bring-up diagnostics for a GPU box: runs each stage against the oracle and prints the numbers
(never asserts), so that one gpurun call yields the full picture.  Output: gpurun_out/diag.log"""
import os
import sys
import time
import traceback

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

from helpers import O, golden_small_case, small_frame_case, to_device  # noqa: E402
from animatable_nerf_b200 import _lib, config, synthetic  # noqa: E402
from animatable_nerf_b200.tpose_nerf_network import Network  # noqa: E402
from animatable_nerf_b200.tpose_renderer import Renderer  # noqa: E402

dev = torch.device('cuda:0')
os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
LOG = open(os.path.join(ROOT, 'gpurun_out', 'diag.log'), 'a')


def say(*a):
    msg = ' '.join(str(x) for x in a)
    print(msg, flush=True)
    LOG.write(msg + '\n')
    LOG.flush()


def md(a, b):
    return float((a.detach().cpu().double() - b.detach().cpu().double()).abs().max())


def section(name, fn):
    say(f'--- {name}')
    try:
        fn()
    except Exception:
        say('EXCEPTION', traceback.format_exc())


def mlp_checks():
    _, _, batch, _ = small_frame_case(voxel=0.05, H=128, W=128, focal=130.0)
    sd = synthetic.make_state_dict(seed=0)
    g = torch.Generator().manual_seed(11)
    n = 4096 + 5
    lo, hi = batch['tbounds'][0, 0], batch['tbounds'][0, 1]
    pts = torch.rand(1, n, 3, generator=g) * (hi - lo) + lo
    vd = torch.nn.functional.normalize(torch.randn(1, n, 3, generator=g), dim=2)
    alpha, rgb = O.nerf_alpha_rgb(sd, pts, vd, batch['latent_index'])
    init = O.sample_blend_weights(pts, batch['tbw'], batch['tbounds'])[:, :24]
    bw = O.neural_blend_weights(sd, pts, init, batch['latent_index'] + 1)
    for swap in (0,):
        for prec in (1, 3):
            try:
                net = Network(config.make_cfg(b200_nerf_precision=prec, b200_bw_precision=prec))
                net.load_state_dict(sd)
                net = net.to(dev)
                ga, gr = net.tpose_human.calculate_alpha_rgb(pts.to(dev), vd.to(dev), batch['latent_index'].to(dev))
                torch.cuda.synchronize()
                say(f'swap={swap} prec={prec} NERF: alpha maxdiff {md(ga, alpha):.3e} (|alpha| max {float(alpha.abs().max()):.3f}) '
                    f'rgb maxdiff {md(gr, rgb):.3e}')
                gb = net.calculate_neural_blend_weights(pts.to(dev), init.to(dev), (batch['latent_index'] + 1).to(dev))
                torch.cuda.synchronize()
                say(f'swap={swap} prec={prec} BW  : bw maxdiff {md(gb, bw):.3e}')
            except Exception:
                say(f'swap={swap} prec={prec} EXCEPTION', traceback.format_exc())
                return


def render_check():
    g, batch, sd = golden_small_case()
    cfg = config.make_cfg(perturb=0.)
    net = Network(cfg)
    net.load_state_dict(sd)
    net = net.to(dev)
    r = Renderer(net, cfg)
    ref = O.render(sd, batch, O.OracleCfg(perturb=0.), return_debug=True)
    dv = r.render_device(to_device(batch, dev), want_bw=True)
    torch.cuda.synchronize()
    na = int(dv['n_active'].item())
    say('n_active gpu', na, 'oracle', int(ref['_debug']['pind'].sum()), 'of', ref['_debug']['pind'].numel())
    if na == int(ref['_debug']['pind'].sum()):
        say('active index equal:', bool(np.array_equal(dv['active_index'][:na].cpu().numpy(), np.nonzero(ref['_debug']['pind'].numpy())[0])))
        say('pbw_all maxdiff', md(dv['pbw_all'][:na], ref['_debug']['pbw_all']), 'tbw_all', md(dv['tbw_all'][:na], ref['_debug']['tbw_all']))
    for k in ('rgb_map', 'acc_map', 'depth_map', 'raw'):
        say(k, 'maxdiff vs oracle', md(dv[k].view(ref[k].shape), ref[k]), 'vs golden', md(dv[k].view(ref[k].shape), torch.from_numpy(g[k])))


def timing():
    frame = synthetic.make_frame(voxel=0.025)
    K, R, T = synthetic.make_camera(frame, 1024, 1024)
    from animatable_nerf_b200 import frontend
    ray_o, ray_d, near, far, mask = frontend.get_rays_within_bounds(1024, 1024, K, R, T, frame['wbounds'], device=dev)
    say('rays in box', ray_o.shape[0])
    batch = synthetic.make_render_batch(frame, ray_o, ray_d, near, far, device=dev)
    sd = synthetic.make_state_dict(seed=0)
    for render_only in (True, False):
        cfg = config.make_cfg(perturb=0., b200_render_only=render_only)
        net = Network(cfg)
        net.load_state_dict(sd)
        net = net.to(dev)
        r = Renderer(net, cfg)
        for it in range(3):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            dv = r.render_device(batch, want_bw=not render_only)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            na = int(dv['n_active'].item())
            say(f'render_only={render_only} iter {it}: {dt * 1e3:.2f} ms, n_active {na} ({na / (ray_o.shape[0] * 64):.3f}), '
                f'{ray_o.shape[0] * 64 / dt / 1e6:.1f} Msamples/s')


if __name__ == '__main__':
    say('=== gpu_diag', time.strftime('%H:%M:%S'), torch.cuda.get_device_name(0))
    which = sys.argv[1:] or ['mlp', 'render', 'timing']
    if 'mlp' in which:
        section('mlp', mlp_checks)
    if 'render' in which:
        section('render', render_check)
    if 'timing' in which:
        section('timing', timing)


def trace():
    """clock64 timeline of one tile of each MLP kernel (block 0, first tile)."""
    _, _, batch, _ = small_frame_case(voxel=0.05, H=128, W=128, focal=130.0)
    sd = synthetic.make_state_dict(seed=0)
    n = int(os.environ.get('ANINERF_TRACE_N', 148 * 128 * 4))
    g = torch.Generator().manual_seed(11)
    lo, hi = batch['tbounds'][0, 0], batch['tbounds'][0, 1]
    pts = (torch.rand(1, n, 3, generator=g) * (hi - lo) + lo).to(dev)
    vd = torch.nn.functional.normalize(torch.randn(1, n, 3, generator=g), dim=2).to(dev)
    init = torch.softmax(torch.randn(1, 24, n, generator=g), dim=1).to(dev)
    net = Network(config.make_cfg())
    net.load_state_dict(sd)
    net = net.to(dev)
    buf = torch.zeros(256, dtype=torch.int64, device=dev)
    for name, fn in (('NERF x1', lambda: net.tpose_human.calculate_alpha_rgb(pts, vd, batch['latent_index'].to(dev))),
                     ('BW x3', lambda: net.calculate_neural_blend_weights(pts, init, (batch['latent_index'] + 1).to(dev)))):
        fn()
        torch.cuda.synchronize()
        buf.zero_()
        _lib.lib().aninerf_debug_set_trace(buf.data_ptr())
        fn()
        torch.cuda.synchronize()
        _lib.lib().aninerf_debug_set_trace(None)
        t = buf.cpu().numpy()
        starts = [int(x) for x in t[160:250] if x]
        say(f'{name}: tile periods of block 0 (cycles):', [b - a for a, b in zip(starts[:-1], starts[1:])])
        t0 = t[0]
        say(f'{name}: tile start 0, PE done {t[1] - t0}')
        for l in range(9):
            for ts in range(2):
                b = 8 + 16 * l + 8 * ts
                if t[b + 5] == 0:
                    continue
                say(f'  L{l} slot{ts}: mma wait_a {t[b + 4] - t0} woke {t[b + 5] - t0} issued {t[b + 6] - t0} | rows wait {t[b] - t0} '
                    f'woke {t[b + 1] - t0} done {t[b + 2] - t0 if t[b + 2] else 0}  || issue {t[b + 6] - t[b + 5]} '
                    f'epilogue {t[b + 2] - t[b + 1] if t[b + 2] else 0}')


def print_trace(name, t):
    starts = [int(x) for x in t[160:250] if x]
    per = [b - a for a, b in zip(starts[:-1], starts[1:])]
    say(f'{name}: {len(starts)} tiles on block 0, mean period {np.mean(per) if per else 0:.0f} cycles; periods {per[:12]} ...')
    t0 = t[0]
    say(f'{name}: tile start 0, PE done {t[1] - t0}, smpl gather {t[4] - t0}..{t[5] - t0}, tile done {t[3] - t0}')
    for l in range(9):
        b = 8 + 16 * l
        if t[b + 5] == 0:
            continue
        rel = lambda x: int(x - t0) if x else 0
        say(f'  L{l}: mma start {rel(t[b + 4])} first stage {rel(t[b + 5])} got half {rel(t[b + 3])} quarters {[rel(x) for x in t[b + 12:b + 16]]} '
            f'issued {rel(t[b + 6])} | rows wait {rel(t[b])} woke {rel(t[b + 1])} published half {rel(t[b + 7])} quarters '
            f'{[rel(x) for x in t[b + 8:b + 12]]} done {rel(t[b + 2])}')


def trace_frame():
    """clock64 timeline of a steady-state tile of each MLP kernel inside the bench frame (volume gather + LBS head)."""
    import bench
    from animatable_nerf_b200 import frontend
    frame, cam, sd = bench.build_workload(1024)
    K, R, T = cam
    ray_o, ray_d, near, far, mask = frontend.get_rays_within_bounds(1024, 1024, K, R, T, frame['wbounds'], device=dev)
    cfg = config.make_cfg(perturb=0., b200_render_only=True)
    net = Network(cfg)
    net.load_state_dict(sd)
    net = net.to(dev).eval()
    r = Renderer(net, cfg)
    batch = synthetic.make_render_batch(frame, ray_o, ray_d, near, far, device=dev)
    r.render_device(batch, want_bw=False)
    torch.cuda.synchronize()
    buf = torch.zeros(512, dtype=torch.int64, device=dev)
    for name, field in (('BW x3 (frame)', 0), ('NERF x1 (frame)', 2)):
        os.environ['ANINERF_TRACE_FIELD'] = str(field)
        buf.zero_()
        _lib.lib().aninerf_debug_set_trace(buf.data_ptr())
        r.render_device(batch, want_bw=False)
        torch.cuda.synchronize()
        _lib.lib().aninerf_debug_set_trace(None)
        t = buf.cpu().numpy()
        print_trace(name, t)
        say(f'{name} weight steps (slot free, copies issued -> issuer has it; relative to tile start): ' + ' | '.join(
            f's{i}: {int(t[416 + 2 * i] - t[0])} {int(t[417 + 2 * i] - t[0])} -> {int(t[448 + i] - t[0])}' for i in range(16) if t[448 + i]))
        tp = t[256:]
        say(f'{name} PEER CTA rows (own clock, relative to its tile start): ' + ' | '.join(
            f'L{l} woke {tp[8 + 16 * l + 1] - tp[0]} done {tp[8 + 16 * l + 2] - tp[0] if tp[8 + 16 * l + 2] else 0}' for l in range(9)))


def mlp_time():
    """Device time of the two MLP kernels inside the bench frame (the library's stage events), no tracing: for A/B comparisons of
    two builds of the library on ONE box (tools/ab.sh)."""
    import ctypes as C
    import bench
    from animatable_nerf_b200 import frontend
    frame, cam, sd = bench.build_workload(1024)
    K, R, T = cam
    ray_o, ray_d, near, far, mask = frontend.get_rays_within_bounds(1024, 1024, K, R, T, frame['wbounds'], device=dev)
    cfg = config.make_cfg(perturb=0., b200_render_only=True)
    net = Network(cfg)
    net.load_state_dict(sd)
    r = Renderer(net.to(dev).eval(), cfg)
    batch = synthetic.make_render_batch(frame, ray_o, ray_d, near, far, device=dev)
    L = _lib.lib()
    for _ in range(5):
        r.render_device(batch, want_bw=False)
    torch.cuda.synchronize()
    ms_buf, calls_buf = (C.c_double * 9)(), (C.c_int64 * 9)()
    L.aninerf_profile_read(ms_buf, calls_buf, 1)
    L.aninerf_profile_enable(1)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(40):
        r.render_device(batch, want_bw=False)
    b.record()
    torch.cuda.synchronize()
    L.aninerf_profile_enable(0)
    L.aninerf_profile_read(ms_buf, calls_buf, 1)
    st = {name: ms_buf[i] / max(1, calls_buf[i]) for i, name in enumerate(_lib.STAGES) if calls_buf[i]}
    say(f"mlp_time: frame {a.elapsed_time(b) / 40:.4f} ms | bw_field_posed {st.get('bw_field_posed', 0):.4f} ms | nerf_field {st.get('nerf_field', 0):.4f} ms")


if 'mlp_time' in sys.argv[1:]:
    section('mlp_time', mlp_time)
if 'trace' in sys.argv[1:]:
    section('trace', trace)
if 'trace_frame' in sys.argv[1:]:
    section('trace_frame', trace_frame)
