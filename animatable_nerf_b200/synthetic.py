"""Deterministic synthetic frames with the reference's batch schema.

The licensed ZJU-MoCap / H36M data and the SMPL model files are absent, so benchmarks and
parity tests run on a synthetic SMPL-like body: 24-joint kinematic tree,
6890 vertices on capsules around the bones, 4-bone skinning weights, a random pose, blend-weight
volumes built with the recipe of `tools/custom_dataset/prepare_blend_weights.py`:156-211 (grid step
0.025 over the vertex box +-0.05, per voxel the weights of the nearest vertex and the distance to
it as channel 24, x-major channels-last), and a pinhole camera.  The batch keys/dtypes/shapes are
the ones `lib/datasets/tpose_dataset.py`:236-277 emits (leading batch dim 1 after collation).
"""
from __future__ import annotations

import numpy as np
import torch
from scipy.spatial import cKDTree

from .host_geometry import SMPL_PARENTS, axis_angle_to_matrix, bone_transforms, bounds_of, look_at_camera

# approximate SMPL neutral T-pose joints (metres); any plausible skeleton works
TPOSE_JOINTS = np.array([
    (-0.002, -0.223, 0.028), (0.070, -0.314, 0.024), (-0.068, -0.315, 0.021), (-0.004, -0.114, 0.002),
    (0.102, -0.690, 0.017), (-0.108, -0.696, 0.015), (0.002, 0.021, 0.003), (0.088, -1.088, -0.027),
    (-0.092, -1.094, -0.027), (0.003, 0.074, 0.028), (0.115, -1.144, 0.093), (-0.117, -1.143, 0.096),
    (0.000, 0.288, -0.015), (0.081, 0.196, -0.006), (-0.079, 0.193, -0.011), (0.005, 0.353, 0.037),
    (0.172, 0.226, -0.015), (-0.172, 0.225, -0.014), (0.432, 0.213, -0.042), (-0.429, 0.213, -0.042),
    (0.681, 0.222, -0.044), (-0.680, 0.222, -0.044), (0.765, 0.214, -0.059), (-0.769, 0.214, -0.057)])

# capsule radius of the bone ending at joint k
_BONE_RADIUS = np.array([0.13, 0.085, 0.085, 0.13, 0.065, 0.065, 0.125, 0.05, 0.05, 0.12, 0.04, 0.04,
                         0.06, 0.08, 0.08, 0.09, 0.055, 0.055, 0.045, 0.045, 0.035, 0.035, 0.035, 0.035])
N_VERTS = 6890


def _segment_distance(p, a, b):
    """distance from points p (n,3) to segments a->b (m,3) -> (n,m)"""
    ab = b - a
    t = np.einsum('nmk,mk->nm', p[:, None] - a[None], ab) / np.maximum((ab * ab).sum(1), 1e-12)
    t = np.clip(t, 0.0, 1.0)
    q = a[None] + t[..., None] * ab[None]
    return np.linalg.norm(p[:, None] - q, axis=2)


def make_body(seed: int = 1):
    """T-pose vertices (6890,3) f32, skinning weights (6890,24) f32, joints (24,3) f64."""
    rng = np.random.RandomState(seed)
    J = TPOSE_JOINTS.copy()
    child = np.arange(1, 24)
    a, b = J[SMPL_PARENTS[child]], J[child]
    length = np.linalg.norm(b - a, axis=1) + 2 * _BONE_RADIUS[child]
    share = length * _BONE_RADIUS[child]
    counts = np.floor(share / share.sum() * N_VERTS).astype(int)
    counts[0] += N_VERTS - counts.sum()
    verts = []
    for i, k in enumerate(child):
        n = counts[i]
        axis = b[i] - a[i]
        L = np.linalg.norm(axis)
        axis = axis / L
        tmp = np.array([1.0, 0, 0]) if abs(axis[0]) < 0.9 else np.array([0, 1.0, 0])
        u = np.cross(axis, tmp)
        u /= np.linalg.norm(u)
        v = np.cross(axis, u)
        s = rng.uniform(-_BONE_RADIUS[k], L + _BONE_RADIUS[k], n)
        phi = rng.uniform(0, 2 * np.pi, n)
        r = np.full(n, _BONE_RADIUS[k])
        over = np.maximum(np.maximum(-s, s - L), 0.0)          # hemispherical caps
        r = np.sqrt(np.maximum(r * r - over * over, 1e-6))
        verts.append(a[i] + s[:, None] * axis + r[:, None] * (np.cos(phi)[:, None] * u + np.sin(phi)[:, None] * v))
    verts = np.concatenate(verts).astype(np.float32)
    d = _segment_distance(verts.astype(np.float64), a, b)       # (V,23) -> bone k=child index
    w = np.zeros((N_VERTS, 24))
    near4 = np.argsort(d, axis=1)[:, :4]
    rows = np.arange(N_VERTS)[:, None]
    w[rows, child[near4]] = 1.0 / np.maximum(d[rows, near4], 1e-3) ** 2
    w /= w.sum(1, keepdims=True)
    return verts, w.astype(np.float32), J


def blend_weight_volume(verts: np.ndarray, weights: np.ndarray, voxel: float = 0.025) -> np.ndarray:
    """(X,Y,Z,25) f32: nearest-vertex skinning weights + distance (channel 24)."""
    verts = verts.astype(np.float64)
    lo = verts.min(0) - 0.05
    hi = verts.max(0) + 0.05
    axes = [np.arange(lo[i], hi[i] + voxel, voxel) for i in range(3)]
    grid = np.stack(np.meshgrid(*axes, indexing='ij'), axis=-1)
    dist, idx = cKDTree(verts).query(grid.reshape(-1, 3))
    vol = np.concatenate([weights[idx], dist[:, None].astype(np.float32)], axis=1)
    return vol.reshape(*grid.shape[:3], 25).astype(np.float32)


def make_frame(pose_seed: int = 2, body_seed: int = 1, voxel: float = 0.025, latent_index: int = 0):
    """Per-frame SMPL-side inputs (numpy): A, R, Th, volumes, bounds."""
    tverts, w, J = make_body(body_seed)
    rng = np.random.RandomState(pose_seed)
    poses = rng.normal(0, 0.2, (24, 3))
    poses[0] = 0
    Rh = rng.normal(0, 0.3, 3)
    Th = (np.array([0, 0, 1.0]) + rng.normal(0, 0.1, 3)).astype(np.float32)
    A = bone_transforms(poses, J)
    M = np.einsum('vk,kij->vij', w.astype(np.float64), A.astype(np.float64))
    pverts = (np.einsum('vij,vj->vi', M[:, :3, :3], tverts.astype(np.float64)) + M[:, :3, 3]).astype(np.float32)
    R = axis_angle_to_matrix(Rh).astype(np.float32)
    wverts = (pverts.astype(np.float64) @ R.T.astype(np.float64) + Th).astype(np.float32)
    return {
        'A': A, 'R': R, 'Th': Th.reshape(1, 3), 'poses': poses, 'joints': J,
        'pbw': blend_weight_volume(pverts, w, voxel), 'tbw': blend_weight_volume(tverts, w, voxel),
        'pbounds': bounds_of(pverts), 'wbounds': bounds_of(wverts), 'tbounds': bounds_of(tverts),
        'latent_index': np.int64(latent_index), 'bw_latent_index': np.int64(latent_index),
        'wverts': wverts,
    }


def make_camera(frame: dict, H: int = 1024, W: int = 1024, focal: float = 1070.0, distance: float = 3.0,
                azimuth: float = 0.0):
    center = frame['wbounds'].astype(np.float64).mean(0)
    R, T = look_at_camera(center, distance, azimuth)
    K = np.array([[focal, 0, W / 2.0], [0, focal, H / 2.0], [0, 0, 1.0]])
    return K, R, T


def make_camera_rig(frame: dict, n_views: int = 4, distance: float = 3.0):
    """World->camera matrices (n_views,4,4) float64 of a ring of cameras around the body (the capture rig of a sweep)."""
    rig = np.zeros((n_views, 4, 4))
    for i in range(n_views):
        _, R, T = make_camera(frame, 64, 64, distance=distance, azimuth=2 * np.pi * i / n_views)
        rig[i, :3, :3], rig[i, :3, 3], rig[i, 3, 3] = R, np.asarray(T).ravel(), 1.0
    return rig


FRAME_KEYS = ('A', 'R', 'Th', 'pbw', 'tbw', 'pbounds', 'wbounds', 'tbounds', 'latent_index', 'bw_latent_index')


def collate_frame(frame: dict, device='cpu') -> dict:
    """numpy frame -> torch batch with the leading batch dim of 1 (default_collate, batch_size 1)."""
    out = {}
    for k in FRAME_KEYS:
        v = frame[k]
        t = torch.as_tensor(np.asarray(v))
        out[k] = t[None].to(device)
    return out


def make_render_batch(frame: dict, ray_o, ray_d, near, far, device='cpu') -> dict:
    """Full `Renderer.render(batch)` input: frame keys + per-ray keys, all with batch dim 1."""
    b = collate_frame(frame, device)
    for k, v in (('ray_o', ray_o), ('ray_d', ray_d), ('near', near), ('far', far)):
        b[k] = torch.as_tensor(np.asarray(v) if not torch.is_tensor(v) else v)[None].to(device)
    n = b['near'].shape[1]
    b['occupancy'] = torch.ones(1, n, dtype=torch.uint8, device=device)
    return b


# ---------------------------------------------------------------------------------------------
# random-init network weights with the reference's checkpoint layout
# ---------------------------------------------------------------------------------------------
def make_state_dict(seed: int = 0, num_train_frame: int = 60, num_eval_frame: int = 0, gain: float = 1.0):
    """`state_dict` with the key names / shapes of the reference `Network`
    (lib/networks/bw_deform/tpose_nerf_network.py:12-38, 219-239, 279-294).
    Values come from numpy's RandomState (stable across torch versions): conv weights and biases
    U(-1/sqrt(fan_in), 1/sqrt(fan_in)) * gain -- the scale of PyTorch's default Conv1d init --
    embeddings N(0,1).  `num_eval_frame > 0` adds the stage-2 `novel_pose_bw.*` field."""
    rng = np.random.RandomState(seed)
    sd = {}

    def conv(name, n_out, n_in):
        bound = gain / np.sqrt(n_in)
        sd[name + '.weight'] = torch.from_numpy(rng.uniform(-bound, bound, (n_out, n_in, 1)).astype(np.float32))
        sd[name + '.bias'] = torch.from_numpy(rng.uniform(-bound, bound, (n_out,)).astype(np.float32))

    def trunk(prefix, input_ch):
        for i in range(8):
            conv(f'{prefix}.{i}', 256, input_ch if i == 0 else 256 + (input_ch if i == 5 else 0))

    sd['tpose_human.nf_latent.weight'] = torch.from_numpy(rng.normal(0, 1, (num_train_frame, 128)).astype(np.float32))
    trunk('tpose_human.pts_linears', 63)
    conv('tpose_human.alpha_fc', 1, 256)
    conv('tpose_human.feature_fc', 256, 256)
    conv('tpose_human.latent_fc', 256, 384)
    conv('tpose_human.view_fc', 128, 283)
    conv('tpose_human.rgb_fc', 3, 128)
    sd['bw_latent.weight'] = torch.from_numpy(rng.normal(0, 1, (num_train_frame + 1, 128)).astype(np.float32))
    trunk('bw_linears', 191)
    conv('bw_fc', 24, 256)
    if num_eval_frame > 0:
        sd['novel_pose_bw.bw_latent.weight'] = torch.from_numpy(rng.normal(0, 1, (num_eval_frame, 128)).astype(np.float32))
        trunk('novel_pose_bw.bw_linears', 191)
        conv('novel_pose_bw.bw_fc', 24, 256)
    return sd


def make_rays(frame: dict, H: int = 1024, W: int = 1024, focal: float = 1070.0, distance: float = 3.0, azimuth: float = 0.0):
    """Host-side (numpy restatement free) camera for a frame: returns K, R, T for the GPU front end."""
    return make_camera(frame, H, W, focal, distance, azimuth)


def make_train_batch(frame: dict, ray_o, ray_d, near, far, n_rays: int = 1024, ray_seed: int = 3, rgb_seed: int = 4,
                     jitter_seed: int = 5, device='cpu'):
    """A training batch as lib/datasets/tpose_dataset.py:236-277 collates it: `n_rays` seeded
    rays of the box-hitting set, target colours U(0,1), `mask_at_box`, and the stratified jitter `t_rand` drawn on the
    CPU generator as tpose_renderer.py:35 does.  Returns (batch, t_rand (1, n_rays, 64))."""
    rs = np.random.RandomState(ray_seed)
    total = np.asarray(near).shape[0]
    sel = np.sort(rs.choice(total, size=min(n_rays, total), replace=False))
    b = make_render_batch(frame, np.asarray(ray_o)[sel], np.asarray(ray_d)[sel], np.asarray(near)[sel], np.asarray(far)[sel], device=device)
    n = sel.shape[0]
    b['rgb'] = torch.from_numpy(np.random.RandomState(rgb_seed).uniform(0, 1, (1, n, 3)).astype(np.float32)).to(device)
    m = np.ones((1, n), dtype=bool)
    m[0, ::17] = False                       # a few sampled pixels outside the box mask (if_nerf_data_utils.py:283-292)
    b['mask_at_box'] = torch.from_numpy(m).to(device)
    g = torch.Generator().manual_seed(jitter_seed)
    t_rand = torch.rand(1, n, 64, generator=g)
    return b, t_rand


# ---------------------------------------------------------------------------------------------
# training-view silhouettes for the novel-view renderer (tpose_renderer_mmsk)
# ---------------------------------------------------------------------------------------------
def make_silhouettes(frame: dict, n_views: int = 4, H: int = 256, W: int = 256, focal: float = 270.0, distance: float = 3.0,
                     radius: int = 6):
    """`msks (V,H,W) u8`, `Ks (V,3,3) f32`, `RT (V,4,4) f32` as lib/datasets/tpose_novel_view_dataset.py:123-194
    prepares them: per training view the body silhouette, dilated (the reference dilates the mask the same way).
    Here the silhouette is the union of discs of `radius` pixels around the projected posed vertices."""
    v = frame['wverts'].astype(np.float64)
    msks = np.zeros((n_views, H, W), dtype=np.uint8)
    Ks = np.zeros((n_views, 3, 3), dtype=np.float32)
    RT = np.zeros((n_views, 4, 4), dtype=np.float32)
    yy, xx = np.mgrid[-radius:radius + 1, -radius:radius + 1]
    disc = (yy * yy + xx * xx) <= radius * radius
    for i in range(n_views):
        K, R, T = make_camera(frame, H, W, focal=focal, distance=distance, azimuth=2 * np.pi * i / n_views)
        cam = v @ R.T + T.ravel()
        uv = cam @ K.T
        uv = np.rint(uv[:, :2] / uv[:, 2:]).astype(int)
        m = np.zeros((H + 2 * radius, W + 2 * radius), dtype=bool)
        ok = (uv[:, 0] >= 0) & (uv[:, 0] < W) & (uv[:, 1] >= 0) & (uv[:, 1] < H)
        for x, y in uv[ok]:
            m[y:y + 2 * radius + 1, x:x + 2 * radius + 1] |= disc
        msks[i] = m[radius:radius + H, radius:radius + W]
        Ks[i] = K.astype(np.float32)
        RT[i, :3, :3] = R.astype(np.float32)
        RT[i, :3, 3] = T.ravel().astype(np.float32)
        RT[i, 3, 3] = 1
    return msks, Ks, RT


def add_silhouettes(batch: dict, frame: dict, device='cpu', **kw) -> dict:
    msks, Ks, RT = make_silhouettes(frame, **kw)
    b = dict(batch)
    b['msks'] = torch.from_numpy(msks)[None].to(device)
    b['Ks'] = torch.from_numpy(Ks)[None].to(device)
    b['RT'] = torch.from_numpy(RT)[None].to(device)
    b['H'] = torch.tensor([msks.shape[1]])
    b['W'] = torch.tensor([msks.shape[2]])
    return b
