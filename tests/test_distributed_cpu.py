"""CPU, world_size 2 over gloo: the ray-tile gather that the N-GPU render uses (NCCL on the box)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_rays, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from animatable_nerf_b200 import ray_tiles
    full = torch.arange(n_rays * 5, dtype=torch.float32).view(n_rays, 5)
    mine = full[ray_tiles.shard_indices(n_rays, rank, world)]
    got = ray_tiles.gather_maps(mine, n_rays, rank, world)
    q.put((rank, bool(torch.equal(got, full))))
    dist.destroy_process_group()


def test_gather_maps_world2_gloo():
    ctx = mp.get_context('spawn')
    for n_rays in (5000, 2048, 3):
        q = ctx.Queue()
        port = _free_port()
        ps = [ctx.Process(target=_worker, args=(r, 2, port, n_rays, q)) for r in range(2)]
        for p in ps:
            p.start()
        res = [q.get(timeout=120) for _ in ps]
        for p in ps:
            p.join(timeout=60)
        assert sorted(res) == [(0, True), (1, True)]


def _worker_train(rank, world, port, q):
    """allreduce_gradients (the DDP mean of trainer.py:13-19 as one flat all-reduce), view-sweep gather and density-grid
    gather on CPU tensors over gloo."""
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from animatable_nerf_b200 import config, sweep
    from animatable_nerf_b200.tpose_nerf_network import Network
    from animatable_nerf_b200.tpose_trainer import allreduce_gradients
    net = Network(config.make_cfg())
    for i, p in enumerate(net.parameters()):
        p.grad = torch.full_like(p, float(rank + 1) * (i + 1))
    allreduce_gradients(net, world)
    ok = all(torch.allclose(p.grad, torch.full_like(p, 1.5 * (i + 1))) for i, p in enumerate(net.parameters()))
    # sweep gather: view v rendered by rank v % world carries the constant v
    n_views, H, W = 5, 4, 3
    local = {v: (torch.full((H, W, 3), float(v)), None, None) for v in sweep.views_of_rank(n_views, rank, world)}
    stack = sweep.gather_views(local, n_views, H, W, rank, world, torch.device('cpu'))
    ok = ok and all(bool((stack[v] == v).all()) for v in range(n_views))
    q.put((rank, ok))
    dist.destroy_process_group()


def test_gradient_allreduce_and_view_gather_world2_gloo():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker_train, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = [q.get(timeout=180) for _ in ps]
    for p in ps:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]
