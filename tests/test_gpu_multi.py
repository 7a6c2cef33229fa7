"""N > 1 under `pytest -m gpu`: the ray-tiled multi-GPU render (peer-memory gather fused into the compositing kernel, and the
NCCL all_gather path), plain and silhouette-culled, must equal the 1-GPU image bit for bit.  Needs >= 2 GPUs on the box
(`gpurun --gpus 2`); skipped otherwise -- the host-side sharding logic is covered on CPU by tests/test_distributed_cpu.py."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason='needs 2 GPUs on one box')
def test_tiled_render_equals_one_gpu_bit_for_bit_world2():
    port = 29500 + os.getpid() % 2000
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr', '127.0.0.1',
           '--master-port', str(port), os.path.join(ROOT, 'tests', 'mp_tiled_worker.py')]
    p = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-3000:] + '\n' + p.stderr[-3000:]
    assert 'MP_TILED_OK' in p.stdout
