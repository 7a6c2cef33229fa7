"""Run the stand-alone stage kernels (volume sampling, inverse LBS, compositing) a few times on L2-exceeding inputs --
the command ncu wraps for the HBM-bound kernels' `--set full` captures (profiles/rNN_stage_kernels_*.md):

    ncu --set full --clock-control none --import-source on -k regex:'sample_bw_kernel|lbs_kernel|composite_kernel' -c 6 \
        -o gpurun_out/prof_stage -f python tools/stage_kernels.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from animatable_nerf_b200 import _lib, config, frontend  # noqa: E402
from animatable_nerf_b200.tpose_renderer import Renderer  # noqa: E402

dev = torch.device('cuda:0')
frame, cam, sd = bench.build_workload(1024)
K, R, T = cam
ro, rd, near, far, _ = frontend.get_rays_within_bounds(1024, 1024, K, R, T, frame['wbounds'], device=dev)
L = _lib.lib()
nr, S = 65536, 64
tv = torch.linspace(0., 1., steps=S).to(dev)
wp = torch.empty(nr * S, 3, device=dev)
_lib.check(L.aninerf_sample_points(_lib.ptr(ro), _lib.ptr(rd), _lib.ptr(near), _lib.ptr(far), _lib.ptr(tv), None, nr, S, _lib.ptr(wp), None, None,
                                   _lib.stream_ptr(dev)))
pp = torch.empty_like(wp)
Rm, Th = torch.as_tensor(frame['R']).to(dev), torch.as_tensor(frame['Th']).to(dev)
_lib.check(L.aninerf_world_to_pose(_lib.ptr(wp), nr * S, _lib.ptr(Rm), _lib.ptr(Th), _lib.ptr(pp), _lib.stream_ptr(dev)))
out = bench.stage_kernel_rooflines(dev, frame, bench.peaks()['hbm_gbs'], ray_pts=pp)
torch.cuda.synchronize()
for k, v in out.items():
    print(k, f"{v['ms']:.4f} ms  {v['achieved_gbs']:.0f} GB/s  {100 * v['frac']:.1f}% of HBM")
