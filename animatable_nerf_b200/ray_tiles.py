"""Ray-tile sharding of one frame across the GPUs of a box (one process per GPU).

Rays are independent; the only cross-ray coupling in the reference is the
per-2048-ray-chunk argmin/argmax forcing (tpose_nerf_network.py:154, :193-194), so tiles are whole
chunks of the reference's chunk grid, dealt round-robin to the ranks (load balance: neighbouring
chunks cover neighbouring image rows).  Every rank then renders its tiles with the same kernels and
the N-GPU image is bit-identical to the 1-GPU image.  The only collective is the final gather of the
(rgb, acc, depth) tiles -- 20 B/ray -- over NCCL (gloo in the CPU tests).
"""
from __future__ import annotations

import torch
import torch.distributed as dist

CHUNK = 2048
PER_RAY_KEYS = ('ray_o', 'ray_d', 'near', 'far', 'occupancy', 'rgb', 'mask_at_box')


def n_chunks(n_rays: int, chunk: int = CHUNK) -> int:
    return (n_rays + chunk - 1) // chunk


def chunks_of_rank(n_rays: int, rank: int, world: int, chunk: int = CHUNK):
    return list(range(rank, n_chunks(n_rays, chunk), world))


def shard_indices(n_rays: int, rank: int, world: int, chunk: int = CHUNK, device='cpu') -> torch.Tensor:
    """Ray indices (ascending within each chunk, chunks in round-robin order) rendered by `rank`."""
    parts = [torch.arange(c * chunk, min(n_rays, (c + 1) * chunk), device=device) for c in chunks_of_rank(n_rays, rank, world, chunk)]
    return torch.cat(parts) if parts else torch.zeros(0, dtype=torch.long, device=device)


def shard_sizes(n_rays: int, world: int, chunk: int = CHUNK):
    return [sum(min(n_rays, (c + 1) * chunk) - c * chunk for c in chunks_of_rank(n_rays, r, world, chunk)) for r in range(world)]


def shard_batch(batch: dict, rank: int, world: int, chunk: int = CHUNK) -> dict:
    """The rank's slice of a `Renderer.render` batch: per-ray keys are gathered, frame keys are shared."""
    n_rays = batch['ray_o'].shape[1]
    idx = shard_indices(n_rays, rank, world, chunk, device=batch['ray_o'].device)
    out = dict(batch)
    for k in PER_RAY_KEYS:
        if k in batch and torch.is_tensor(batch[k]) and batch[k].dim() >= 2 and batch[k].shape[1] == n_rays:
            out[k] = batch[k].index_select(1, idx)
    return out


_PLANS = {}


def gather_plan(n_rays: int, world: int, chunk: int = CHUNK, device='cpu'):
    """(pad, pos): pad = rows every rank contributes to the all-gather; pos[i] = row of frame ray i in the
    gathered (world*pad, C) buffer.  Cached: a sweep renders many frames with the same ray count."""
    key = (n_rays, world, chunk, str(device))
    if key not in _PLANS:
        sizes = shard_sizes(n_rays, world, chunk)
        pad = max(sizes)
        pos = torch.empty(n_rays, dtype=torch.long)
        for r in range(world):
            idx = shard_indices(n_rays, r, world, chunk)
            pos[idx] = r * pad + torch.arange(sizes[r])
        _PLANS[key] = (pad, pos.to(device))
    return _PLANS[key]


def gather_maps(local: torch.Tensor, n_rays: int, rank: int, world: int, chunk: int = CHUNK, group=None) -> torch.Tensor:
    """All-gather per-ray rows (R_local, C) from every rank and put them back in frame order -> (n_rays, C).
    One collective (NCCL all_gather over NVLink, 20 B/ray for rgb+acc+depth) and one gather kernel."""
    if world == 1:
        return local
    pad, pos = gather_plan(n_rays, world, chunk, local.device)
    C = local.shape[1]
    if local.shape[0] == pad:
        send = local.contiguous()
    else:
        send = torch.zeros(pad, C, dtype=local.dtype, device=local.device)
        send[:local.shape[0]] = local
    recv = torch.empty(world * pad, C, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(recv, send, group=group)
    return recv.index_select(0, pos)


class PeerImage:
    """Frame image buffers in symmetric memory (every rank's buffer peer-mapped into every other rank over NVLink/NVSwitch)
    for the fused compositing + gather of `aninerf_render_rays_tiled`: the compositing kernel of each rank stores its rays'
    (rgb, acc, depth) rows into ALL ranks' buffers at their frame position, then ONE barrier makes the image complete
    everywhere -- no all_gather, no reorder.  Two buffers alternate so that a rank may still read frame k while a faster
    peer already writes frame k+1 (stream-ordered consumers; the barrier of frame k+1 orders everything else)."""

    def __init__(self, capacity_rays: int, rank: int, world: int, device, group=None):
        import torch.distributed._symmetric_memory as symm
        from . import _lib
        self.rank, self.world, self.capacity = rank, world, int(capacity_rays)
        group = group if group is not None else dist.group.WORLD
        self.bufs = [symm.empty(self.capacity * 5, dtype=torch.float32, device=device) for _ in range(2)]
        self.hdls = [symm.rendezvous(b, group) for b in self.bufs]
        self.structs = []
        for h in self.hdls:
            pg = _lib.PeerGather()
            ptrs = list(h.buffer_ptrs)
            for k in range(world):
                pg.maps[k] = ptrs[k]
            pg.world, pg.rank = world, rank
            self.structs.append(pg)
        self.turn = 0

    def begin(self):
        """-> (PeerGather for render_device(peers=...), slot) of the next frame"""
        slot = self.turn
        self.turn ^= 1
        return self.structs[slot], slot

    def finish(self, slot: int, n_rays: int) -> torch.Tensor:
        """Barrier across the ranks on the current stream, then the complete (n_rays, 5) image of this rank's buffer."""
        assert n_rays <= self.capacity
        self.hdls[slot].barrier(0, 10000)
        return self.bufs[slot][:n_rays * 5].view(n_rays, 5)


class PeerVolume:
    """Sharded host->device upload of a frame's replicated volume (`pbw`, 18 MB at 2.5 cm): every rank needs the whole volume, but
    each rank's PCIe link only carries 1/world of it -- rank r copies slice r from (pinned) host memory into its own buffer and
    pushes it into every peer's buffer over NVLink (peer-mapped symmetric memory, the same pattern as the image gather); one
    barrier completes the volume everywhere.  Two buffers alternate (a slow rank may still sample frame k while a fast one
    uploads frame k+1)."""

    def __init__(self, capacity_floats: int, rank: int, world: int, device, group=None):
        import torch.distributed._symmetric_memory as symm
        self.rank, self.world, self.capacity = rank, world, int(capacity_floats)
        group = group if group is not None else dist.group.WORLD
        self.bufs = [symm.empty(self.capacity, dtype=torch.float32, device=device) for _ in range(2)]
        self.hdls = [symm.rendezvous(b, group) for b in self.bufs]
        self.peers = [[h.get_buffer(p, (self.capacity,), torch.float32, 0) for p in range(world)] for h in self.hdls]
        self.turn = 0

    def slice_of(self, n: int, rank: int):
        per = ((n + self.world - 1) // self.world + 3) // 4 * 4
        return min(n, rank * per), min(n, (rank + 1) * per)

    def upload(self, host_vol: torch.Tensor) -> torch.Tensor:
        """host_vol: float32 host tensor (pinned for an asynchronous copy), identical on every rank.  Returns the device copy
        (a view of this rank's symmetric buffer, valid until the upload after next)."""
        n = host_vol.numel()
        assert n <= self.capacity
        slot = self.turn
        self.turn ^= 1
        flat = host_vol.reshape(-1)
        a, b = self.slice_of(n, self.rank)
        mine = self.bufs[slot][a:b]
        if b > a:
            mine.copy_(flat[a:b], non_blocking=True)                       # 1/world of the volume over this GPU's PCIe link
            for p in range(self.world):
                if p != self.rank:
                    self.peers[slot][p][a:b].copy_(mine, non_blocking=True)    # NVLink push into the peer's buffer
        self.hdls[slot].barrier(0, 10000)
        return self.bufs[slot][:n].view(host_vol.shape)
