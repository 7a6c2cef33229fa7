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


SIGMA_EPS = 5e-3     # |sigma_gpu - sigma_oracle| bound of the bf16 NeRF field used by the row-selection check below


def check_selected_rows(renderer, dv, dbg, bw_tol=1e-5, eps=SIGMA_EPS):
    """The `pbw` / `tbw` rows of the contract (alpha > train_th plus the first arg-max row of every 2048-ray chunk,
    tpose_nerf_network.py:192-196) against the oracle, UNCONDITIONALLY:
      * the active sets are bit-identical, so row i of the GPU is row i of the oracle;
      * a row may be selected on one side only if the oracle's masked sigma is within `eps` of train_th (0), or it is the
        forced arg-max of a chunk whose best and runner-up sigma differ by less than `eps` on the oracle side;
      * on the rows both sides select, pbw and tbw agree within `bw_tol`.
    dv: render_device(want_bw=True) result; dbg: the oracle's _debug (needs alpha_ind, sigma_masked, chunk_active).
    Returns (gpu rows, mask of those the oracle selects too, oracle rows, mask of those the GPU selects too, packed pbw, packed tbw,
    number of one-sided rows, number of oracle rows with |sigma| <= eps -- the population the one-sided rows can come from)."""
    n_active = int(dv['n_active'].item())
    sel, n_sel = renderer.select_rows(dv)
    sel = sel[:n_active].bool().cpu()
    want = dbg['alpha_ind'].cpu()
    assert sel.numel() == want.numel()
    assert int(n_sel.item()) == int(sel.sum())
    sig = dbg['sigma_masked'].cpu()
    mism = (sel != want).nonzero().reshape(-1)
    if mism.numel():
        offs = torch.cat([torch.zeros(1, dtype=torch.long), dbg['chunk_active'].cumsum(0)])
        chunk_of = torch.bucketize(mism, offs[1:], right=True)
        for i, c in zip(mism.tolist(), chunk_of.tolist()):
            near_th = abs(float(sig[i])) <= eps
            seg = sig[int(offs[c]):int(offs[c + 1])]
            top = torch.topk(seg, min(2, seg.numel()))[0]
            near_max = float(top[0] - sig[i]) <= eps
            assert near_th or near_max, f'row {i} (chunk {c}): selected on one side only with sigma {float(sig[i])}, chunk max {float(top[0])}'
    both = sel & want
    k = int(sel.sum())
    pbw, tbw = renderer.gather_selected(dv, renderer.select_rows(dv)[0], k)
    rows = sel.nonzero().reshape(-1)
    common_in_gpu = both[rows]                                   # which of the GPU's packed rows the oracle selected too
    ref_rows = want.nonzero().reshape(-1)
    common_in_ref = both[ref_rows]
    n_near = int((sig.abs() <= eps).sum()) + int(dbg['chunk_active'].numel())
    return rows, common_in_gpu, ref_rows, common_in_ref, pbw.cpu(), tbw.cpu(), int(mism.numel()), n_near
