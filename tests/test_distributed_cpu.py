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
