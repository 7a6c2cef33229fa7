"""Probe: does torch symmetric memory (CUDA peer mappings over NVLink) work on the GPU box?  torchrun, 2+ ranks."""
import os
import sys
import time

import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm

rank, world, lr = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(lr)
dev = torch.device('cuda', lr)
dist.init_process_group('nccl', device_id=dev)
try:
    t = symm.empty(1 << 20, dtype=torch.float32, device=dev)
    t.fill_(float(rank))
    hdl = symm.rendezvous(t, dist.group.WORLD)
    print(f'[{rank}] rendezvous ok: ptrs', [hex(p) for p in hdl.buffer_ptrs], 'multicast', hdl.has_multicast_support, flush=True)
    hdl.barrier()
    peer = (rank + 1) % world
    remote = hdl.get_buffer(peer, (1 << 20,), torch.float32)
    remote[rank * 16:(rank + 1) * 16] = 100.0 + rank          # write into the peer's memory
    hdl.barrier()
    torch.cuda.synchronize()
    src = (rank - 1) % world
    print(f'[{rank}] slice written by rank {src}:', t[src * 16:src * 16 + 2].tolist(), 'own fill', t[-1].item(), flush=True)
    # barrier latency
    for _ in range(5):
        hdl.barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(100):
        hdl.barrier()
    b.record()
    torch.cuda.synchronize()
    print(f'[{rank}] barrier {a.elapsed_time(b) * 10:.1f} us each', flush=True)
except Exception as e:  # noqa: BLE001
    import traceback
    print(f'[{rank}] symmetric memory FAILED:', traceback.format_exc(), flush=True)
sys.stdout.flush()
os._exit(0)
