"""GPU-busy time vs wall time of one training iteration (torch.profiler), to tell launch-bound from kernel-bound."""
import sys
import time

import torch

sys.path.insert(0, '.')
import bench
from animatable_nerf_b200 import config, frontend, synthetic
from animatable_nerf_b200.tpose_nerf_network import Network
from animatable_nerf_b200.tpose_trainer import NetworkWrapper, train_iteration

dev = torch.device('cuda:0')
frame, cam, sd = bench.build_workload(1024)
K, R, T = cam
ro, rd, near, far, _ = frontend.get_rays_within_bounds(1024, 1024, K, R, T, frame['wbounds'], device=dev)
cfg = config.make_cfg(perturb=1.)
net = Network(cfg)
net.load_state_dict(sd)
net = net.to(dev).train()
w = NetworkWrapper(net, cfg)
tb, t_rand = synthetic.make_train_batch(frame, ro.cpu().numpy(), rd.cpu().numpy(), near.cpu().numpy(), far.cpu().numpy(), n_rays=1024, device=dev)
opt = torch.optim.Adam(net.parameters(), lr=5e-4)
for _ in range(3):
    train_iteration(w, tb, opt, t_rand=t_rand)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10):
    train_iteration(w, tb, opt, t_rand=t_rand)
torch.cuda.synchronize()
print('wall ms/iter', (time.perf_counter() - t0) * 100)
from torch.profiler import ProfilerActivity, profile
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        train_iteration(w, tb, opt, t_rand=t_rand)
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
print('GPU busy ms/iter', sum(e.device_time for e in ev) / 3e3, 'kernels/iter', len(ev) / 3)
agg = {}
for e in ev:
    a = agg.setdefault(e.name[:50], [0, 0.0])
    a[0] += 1
    a[1] += e.device_time
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:12]:
    print(f'{k:52s} {n / 3:6.1f} {t / 3e3:8.3f} ms')

