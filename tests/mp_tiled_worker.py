"""Worker of tests/test_gpu_multi.py (launched with torch.distributed.run, one process per GPU, NCCL).

Checks, on every rank, that the ray-tiled N-GPU render equals the 1-GPU render of the same kernels BIT FOR BIT:
  1. fused compositing + peer-memory gather (`aninerf_render_rays_tiled` + ray_tiles.PeerImage)  == whole frame on one GPU;
  2. the NCCL all_gather + reorder path (`ray_tiles.gather_maps`)                                 == the same image;
  3. both again with the multi-view silhouette culling of tpose_renderer_mmsk (culled + tiled);
  4. a second frame through the double-buffered peer images (slot reuse).
Prints 'MP_TILED_OK' from rank 0 on success; any assertion kills the job.
"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from animatable_nerf_b200 import config, frontend, ray_tiles, synthetic  # noqa: E402
from animatable_nerf_b200.tpose_nerf_network import Network  # noqa: E402
from animatable_nerf_b200 import tpose_renderer, tpose_renderer_mmsk  # noqa: E402


def maps_of(out):
    return torch.cat([out['rgb_map'], out['acc_map'][:, None], out['depth_map'][:, None]], dim=1)


def main():
    rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
    local = int(os.environ.get('LOCAL_RANK', rank))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist.init_process_group('nccl', device_id=dev)
    size = int(os.environ.get('MP_TILED_SIZE', '384'))
    frame = synthetic.make_frame(pose_seed=2, body_seed=1, voxel=0.04, latent_index=3)
    K, R, T = synthetic.make_camera(frame, size, size, focal=1070.0 * size / 1024.0)
    ro, rd, near, far, _ = frontend.get_rays_within_bounds(size, size, K, R, T, frame['wbounds'], device=dev)
    n = ro.shape[0]
    assert n > 4 * 2048, n                                     # several chunks per rank, ragged tail
    cfg = config.make_cfg(perturb=0., b200_render_only=True)
    net = Network(cfg)
    net.load_state_dict(synthetic.make_state_dict(seed=0))
    net = net.to(dev).eval()
    full = synthetic.make_render_batch(frame, ro, rd, near, far, device=dev)
    full = synthetic.add_silhouettes(full, frame, device=dev)
    peer = ray_tiles.PeerImage(n, rank, world, dev)
    for name, R_cls in (('plain', tpose_renderer.Renderer), ('culled', tpose_renderer_mmsk.Renderer)):
        r = R_cls(net, cfg)
        whole = maps_of(r.render_device(full, want_bw=False))            # every rank renders the whole frame once: the 1-GPU image
        mine = ray_tiles.shard_batch(full, rank, world)
        for frame_no in range(3):                                         # three frames: both peer-image slots are reused
            pg, slot = peer.begin()
            out = r.render_device(mine, want_bw=False, peers=pg)
            img = peer.finish(slot, n)
            torch.cuda.synchronize()
            assert torch.equal(img, whole), f'{name}: peer-gathered image differs from the 1-GPU image (rank {rank}, frame {frame_no})'
        local_maps = maps_of(r.render_device(mine, want_bw=False))
        img2 = ray_tiles.gather_maps(local_maps, n, rank, world)
        assert torch.equal(img2, whole), f'{name}: NCCL-gathered image differs from the 1-GPU image (rank {rank})'
        idx = ray_tiles.shard_indices(n, rank, world, device=dev)
        assert torch.equal(local_maps, whole[idx]), f'{name}: this rank\'s tile differs from its rows of the 1-GPU image'
        if name == 'culled':
            assert int(out['n_active'].item()) > 0
    dist.barrier()
    if rank == 0:
        print('MP_TILED_OK', n, 'rays', world, 'ranks')
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
