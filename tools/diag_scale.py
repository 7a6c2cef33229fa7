"""N-GPU step diagnostics (torchrun): per-step device times of the bench step under different
conditions (gather on/off, L2 flush, per-stage profiling, CUDA-graph replay).  Prints only."""
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
from animatable_nerf_b200 import _lib, config, frontend, ray_tiles, synthetic  # noqa: E402
from animatable_nerf_b200.tpose_nerf_network import Network  # noqa: E402
from animatable_nerf_b200.tpose_renderer import Renderer  # noqa: E402

world = int(os.environ.get('WORLD_SIZE', '1'))
rank = int(os.environ.get('RANK', '0'))
local_rank = int(os.environ.get('LOCAL_RANK', '0'))
torch.cuda.set_device(local_rank)
dev = torch.device('cuda', local_rank)
if world > 1:
    dist.init_process_group('nccl', device_id=dev)
L = _lib.lib()
frame, cam, sd = bench.build_workload(1024)
K, R, T = cam
ray_o, ray_d, near, far, mask = frontend.get_rays_within_bounds(1024, 1024, K, R, T, frame['wbounds'], device=dev)
n_rays = ray_o.shape[0]
cfg = config.make_cfg(perturb=0., b200_render_only=True)
net = Network(cfg)
net.load_state_dict(sd)
net = net.to(dev).eval()
renderer = Renderer(net, cfg)
full = synthetic.make_render_batch(frame, ray_o, ray_d, near, far, device=dev)
mine = ray_tiles.shard_batch(full, rank, world)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def say(*a):
    print(f'[rank {rank}]', *a, flush=True)


def render_only():
    return renderer.render_device(mine, want_bw=False)


def render_gather():
    out = renderer.render_device(mine, want_bw=False)
    maps = torch.cat([out['rgb_map'], out['acc_map'][:, None], out['depth_map'][:, None]], dim=1)
    return ray_tiles.gather_maps(maps, n_rays, rank, world)


def run(name, fn, steps=10, do_flush=False, profile=False, sync_each=False):
    L.aninerf_profile_enable(1 if profile else 0)
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    evs, cpu = [], []
    for _ in range(steps):
        if do_flush:
            flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        a.record()
        r = fn()
        b.record()
        cpu.append((time.perf_counter() - t0) * 1e3)
        if sync_each:
            torch.cuda.synchronize()
        evs.append((a, b))
        del r
    torch.cuda.synchronize()
    L.aninerf_profile_enable(0)
    ms = [a.elapsed_time(b) for a, b in evs]
    say(f'{name:40s} gpu ms/step: ' + ' '.join(f'{x:.2f}' for x in ms) + ' | cpu ms/step: ' + ' '.join(f'{x:.2f}' for x in cpu))


run('render only', render_only)
run('render only, sync each', render_only, sync_each=True)
run('render only + flush', render_only, do_flush=True)
run('render + gather', render_gather)
run('render + gather, sync each', render_gather, sync_each=True)
run('render + gather + flush', render_gather, do_flush=True)
run('render + gather + flush + profile', render_gather, do_flush=True, profile=True)

# CUDA-graph replay of the whole step
try:
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            render_gather()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        img = render_gather()
    torch.cuda.synchronize()
    ref = render_gather()
    g.replay()
    torch.cuda.synchronize()
    say('graph replay equals eager:', bool(torch.equal(ref, img)))
    run('graph replay', lambda: g.replay())
    run('graph replay + flush', lambda: g.replay(), do_flush=True)
except Exception as e:  # noqa: BLE001
    import traceback
    say('graph capture failed:', traceback.format_exc())
if world > 1:
    dist.destroy_process_group()
