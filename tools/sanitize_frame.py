"""A 2-chunk frame through every tcgen05 kernel instantiation and the training GEMM -- the command compute-sanitizer wraps
(memcheck / racecheck / synccheck; logs summarised under profiles/):

    compute-sanitizer --tool racecheck python tools/sanitize_frame.py
Covers mlp_kernel<3,0,2,1> (bf16x3 blend-weight field), <1,0,2,2> (single-pass blend-weight field), <1,1,2,2> (NeRF field),
<3,1,2,1> (bf16x3 NeRF field), gemm_x3_kernel (one training step), and the front-end / compositing kernels.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from helpers import golden_small_case, to_device  # noqa: E402
from animatable_nerf_b200 import config  # noqa: E402
from animatable_nerf_b200.tpose_nerf_network import Network  # noqa: E402
from animatable_nerf_b200.tpose_renderer import Renderer  # noqa: E402

dev = torch.device('cuda:0')
g, batch, sd = golden_small_case()                     # 2175 rays = one full 2048-ray chunk + a ragged one
b = to_device(batch, dev)
which = sys.argv[1:] or ['render', 'train']
if 'render' in which:
    for bw_p, nf_p in ((3, 1), (1, 3)):
        cfg = config.make_cfg(perturb=0., b200_bw_precision=bw_p, b200_nerf_precision=nf_p)
        net = Network(cfg)
        net.load_state_dict(sd)
        net = net.to(dev).eval()
        r = Renderer(net, cfg)
        for want_bw in (True, False):
            out = r.render_device(b, want_bw=want_bw)
            torch.cuda.synchronize()
            print('render bw_precision', bw_p, 'nerf_precision', nf_p, 'want_bw', want_bw, 'n_active', int(out['n_active'].item()),
                  'acc sum', float(out['acc_map'].sum()))
if 'train' in which:
    import numpy as np
    from helpers import load_golden
    from animatable_nerf_b200.tpose_trainer import NetworkWrapper, train_iteration
    gt = load_golden('train_step_small.npz')
    tb = dict(batch)
    for k in ('ray_o', 'ray_d', 'near', 'far', 'rgb', 'mask_at_box'):
        tb[k] = torch.from_numpy(gt[k])[None]
    tb['occupancy'] = torch.ones(1, tb['near'].shape[1], dtype=torch.uint8)
    cfg = config.make_cfg(perturb=1.)
    net = Network(cfg)
    net.load_state_dict(sd)
    net = net.to(dev).train()
    w = NetworkWrapper(net, cfg)
    opt = torch.optim.SGD(net.parameters(), lr=0.0)
    _, stats = train_iteration(w, to_device(tb, dev), opt, t_rand=torch.from_numpy(gt['t_rand']))
    torch.cuda.synchronize()
    print('train step loss', float(stats['loss']))
print('SANITIZE_FRAME_DONE')
