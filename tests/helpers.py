"""Shared test helpers: seeded synthetic cases and the oracle side of each comparison."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from animatable_nerf_b200 import synthetic  # noqa: E402
from oracle import aninerf_oracle as O  # noqa: E402

GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def golden_small_case():
    """The tests/golden/render_small.npz case: frame + rays + state dict (regenerated from seeds and
    checked against the digest stored with the reference outputs)."""
    g = load_golden('render_small.npz')
    frame = {k: g['frame_' + k] for k in synthetic.FRAME_KEYS}
    batch = synthetic.make_render_batch(frame, g['ray_o'], g['ray_d'], g['near'], g['far'])
    sd = synthetic.make_state_dict(seed=int(g['sd_seed']))
    return g, batch, sd


def to_device(batch, device):
    return {k: (v.to(device) if torch.is_tensor(v) else v) for k, v in batch.items()}


def small_frame_case(voxel=0.05, H=128, W=128, focal=130.0, pose_seed=2, latent_index=3, n_rays=None):
    """A seeded synthetic frame + the box-hitting rays of a small camera (oracle stage 1)."""
    frame = synthetic.make_frame(pose_seed=pose_seed, body_seed=1, voxel=voxel, latent_index=latent_index)
    K, R, T = synthetic.make_camera(frame, H, W, focal=focal)
    ray_o, ray_d, near, far, mask = O.get_rays_within_bounds(H, W, K, R, T, frame['wbounds'])
    if n_rays is not None:
        ray_o, ray_d, near, far = ray_o[:n_rays], ray_d[:n_rays], near[:n_rays], far[:n_rays]
    batch = synthetic.make_render_batch(frame, ray_o, ray_d, near, far)
    return frame, (K, R, T), batch, mask
