"""CPU oracle for Animatable-NeRF's per-ray render hot path.

TEST INFRASTRUCTURE ONLY.  This is a restatement (not a copy) of the reference
algorithm in plain numpy / torch-fp32-on-CPU, written functionally over a
`state_dict` with the reference's checkpoint key names.  Every function cites
the reference file:line (relative to /root/reference) it follows.

Parity status: the reference ships NO tests / golden vectors / fixtures
(SURVEY.md section 4), so this oracle is pinned the other way the task allows:
`oracle/validate_against_reference.py` imports the unmodified reference in the
build container, checks every function here against it (bit-equal on CPU), and
writes the committed fixtures under `tests/golden/`.  `tests/test_oracle_golden.py`
re-checks the oracle against those fixtures wherever the tests run.

Arithmetic = "reference code x installed torch (2.11 CPU)".  The ops used are
the same ATen ops the reference calls (conv1d 1x1, grid_sample, inverse,
softmax, cumprod, matmul/bmm) so results are bit-equal on the same machine.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

N_BONES = 24


# ----------------------------------------------------------------------------
# configuration snapshot (reference: lib/config/config.py:9-137 +
# configs/aninerf_s9p.yaml:58-72)
# ----------------------------------------------------------------------------
class OracleCfg:
    N_samples = 64
    perturb = 0.0
    norm_th = 0.05
    train_th = 0.0
    white_bkgd = False
    xyz_res = 10
    view_res = 4
    box_padding = 0.05
    test_novel_pose = False
    chunk = 2048          # tpose_renderer.py:170

    def __init__(self, **kw):
        for k, v in kw.items():
            if not hasattr(OracleCfg, k):
                raise KeyError(k)
            setattr(self, k, v)


# ----------------------------------------------------------------------------
# stage 1: rays, box intersection (numpy, fp64)  -- if_nerf_data_utils.py
# ----------------------------------------------------------------------------
def get_rays(H, W, K, R, T):
    """if_nerf_data_utils.py:64-89.  o = -R^T T ; d = normalize((K^-1 [u,v,1] - T) R - o)."""
    T = np.asarray(T)
    rays_o = -np.dot(R.T, T).ravel()
    u, v = np.meshgrid(np.arange(W, dtype=np.float32), np.arange(H, dtype=np.float32), indexing='xy')
    pix = np.stack([u, v, np.ones_like(u)], axis=2)
    cam = np.dot(pix, np.linalg.inv(K).T)
    world = np.dot(cam - T.ravel(), R)
    rays_d = world - rays_o[None, None]
    rays_d = rays_d / np.linalg.norm(rays_d, axis=2, keepdims=True)
    rays_o = np.broadcast_to(rays_o, rays_d.shape)
    return rays_o, rays_d


def get_near_far(bounds, ray_o, ray_d):
    """if_nerf_data_utils.py:156-196.  Six slab planes, keep rays with exactly two face hits."""
    b = bounds + np.array([-0.01, 0.01])[:, None]          # float64 promotion (:168)
    num = b[None] - ray_o[:, None]
    t = (num / ray_d[:, None]).reshape(-1, 6)              # plane order min_xyz, max_xyz
    p = t[..., None] * ray_d[:, None] + ray_o[:, None]
    lo_x, lo_y, lo_z, hi_x, hi_y, hi_z = b.ravel()
    eps = 1e-6
    hit = (p[..., 0] >= (lo_x - eps)) * (p[..., 0] <= (hi_x + eps)) * \
          (p[..., 1] >= (lo_y - eps)) * (p[..., 1] <= (hi_y + eps)) * \
          (p[..., 2] >= (lo_z - eps)) * (p[..., 2] <= (hi_z + eps))
    mask_at_box = hit.sum(-1) == 2
    pair = p[mask_at_box][hit[mask_at_box]].reshape(-1, 2, 3)
    o = ray_o[mask_at_box]
    d = ray_d[mask_at_box]
    nd = np.linalg.norm(d, axis=1)
    d0 = np.linalg.norm(pair[:, 0] - o, axis=1) / nd
    d1 = np.linalg.norm(pair[:, 1] - o, axis=1) / nd
    return np.minimum(d0, d1), np.maximum(d0, d1), mask_at_box


def get_rays_within_bounds(H, W, K, R, T, bounds):
    """if_nerf_data_utils.py:310-339."""
    ray_o, ray_d = get_rays(H, W, K, R, T)
    ray_o = ray_o.reshape(-1, 3).astype(np.float32)
    ray_d = ray_d.reshape(-1, 3).astype(np.float32)
    near, far, mask = get_near_far(bounds, ray_o, ray_d)
    return (ray_o[mask], ray_d[mask], near.astype(np.float32), far.astype(np.float32),
            mask.reshape(H, W))


def get_bounds(xyz, box_padding=0.05):
    """if_nerf_data_utils.py:566-579."""
    lo = np.min(xyz, axis=0) - box_padding
    hi = np.max(xyz, axis=0) + box_padding
    return np.stack([lo, hi], axis=0).astype(np.float32)


def batch_rodrigues(poses):
    """if_nerf_data_utils.py:392-411."""
    n = poses.shape[0]
    angle = np.linalg.norm(poses + 1e-8, axis=1, keepdims=True)
    axis = poses / angle
    c = np.cos(angle)[:, None]
    s = np.sin(angle)[:, None]
    rx, ry, rz = np.split(axis, 3, axis=1)
    z = np.zeros([n, 1])
    Kx = np.concatenate([z, -rz, ry, rz, z, -rx, -ry, rx, z], axis=1).reshape([n, 3, 3])
    return np.eye(3)[None] + s * Kx + (1 - c) * np.matmul(Kx, Kx)


def get_rigid_transformation(poses, joints, parents):
    """if_nerf_data_utils.py:414-458.  A_k = G_k(theta) * G_k(0)^-1, float32 (24,4,4)."""
    rot = batch_rodrigues(poses)
    rel = joints.copy()
    rel[1:] -= joints[parents[1:]]
    M = np.concatenate([rot, rel[..., None]], axis=2)
    pad = np.zeros([24, 1, 4])
    pad[..., 3] = 1
    M = np.concatenate([M, pad], axis=1)
    chain = [M[0]]
    for i in range(1, parents.shape[0]):
        chain.append(np.dot(chain[parents[i]], M[i]))
    G = np.stack(chain, axis=0)
    jh = np.concatenate([joints, np.zeros([24, 1])], axis=1)
    G[..., 3] = G[..., 3] - np.sum(G * jh[:, None], axis=2)
    return G.astype(np.float32)


# ----------------------------------------------------------------------------
# stage 1b: stratified sample points  -- tpose_renderer.py:14-69
# ----------------------------------------------------------------------------
def sample_points(ray_o, ray_d, near, far, n_samples=64, t_rand=None):
    """tpose_renderer.py:26-38.  ray_o/ray_d (B,R,3), near/far (B,R).
    t_rand (B,R,S) replaces torch.rand when the jitter branch is wanted."""
    t = torch.linspace(0., 1., steps=n_samples).to(near)
    z = near[..., None] * (1. - t) + far[..., None] * t
    if t_rand is not None:
        mids = .5 * (z[..., 1:] + z[..., :-1])
        upper = torch.cat([mids, z[..., -1:]], -1)
        lower = torch.cat([z[..., :1], mids], -1)
        z = lower + (upper - lower) * t_rand
    pts = ray_o[:, :, None] + ray_d[:, :, None] * z[..., None]
    return pts, z


def sample_dists(z):
    """tpose_renderer.py:63-66: diff of z, last interval duplicated."""
    d = z[..., 1:] - z[..., :-1]
    return torch.cat([d, d[..., -1:]], dim=2)


# ----------------------------------------------------------------------------
# stage 2: blend-weight volume, skinning  -- blend_utils.py
# ----------------------------------------------------------------------------
def world_to_pose(wpts, Rh, Th):
    """blend_utils.py:6-16: (p - Th) @ Rh."""
    return torch.matmul(wpts - Th, Rh)


def sample_blend_weights(pts, vol, bounds):
    """blend_utils.py:119-149.  pts (B,m,3), vol (B,X,Y,Z,25), bounds (B,2,3) -> (B,25,m)."""
    lo = bounds[:, 0]
    hi = bounds[:, 1]
    ext = hi[:, None] - lo[:, None]
    g = (pts.clone() - lo[:, None]) / ext
    g = g * 2 - 1
    g = g[..., [2, 1, 0]]
    v = vol.permute(0, 4, 1, 2, 3)
    out = F.grid_sample(v, g[:, None, None], padding_mode='border', align_corners=True)
    return out[:, :, 0, 0]


def inverse_lbs(ppts, bw, A):
    """blend_utils.py:41-59.  ppts (B,m,3), bw (B,24,m), A (B,24,4,4) -> canonical (B,m,3)."""
    B = ppts.shape[0]
    M = torch.bmm(bw.permute(0, 2, 1), A.view(B, 24, -1)).view(B, -1, 4, 4)
    q = ppts - M[..., :3, 3]
    Rinv = torch.inverse(M[..., :3, :3])
    return torch.sum(Rinv * q[:, :, None], dim=3)


def forward_lbs(tpts, bw, A):
    """blend_utils.py:77-90."""
    B = tpts.shape[0]
    M = torch.bmm(bw.permute(0, 2, 1), A.view(B, 24, -1)).view(B, -1, 4, 4)
    return torch.sum(M[..., :3, :3] * tpts[:, :, None], dim=3) + M[..., :3, 3]


def sample_blend_closest_points(src, ref, values, K=5, exp=1e-8):
    """lib/utils/sample_utils.py:323-349 (the extended networks' replacement of the volume lookup).  `knn_points` comes from
    pytorch3d, which is absent here (SURVEY 8c stubs it): restated as the brute force its documentation defines -- squared
    Euclidean distances, the K smallest in ascending order -- PARITY UNPINNED for that third-party piece; everything after it
    follows the reference line by line.  src (B,n,3), ref (B,V,3), values (B,V,24) -> (B,n,24), (B,n,1)."""
    n_batch, n_points, _ = src.shape
    diff = src[:, :, None, :] - ref[:, None, :, :]
    d2 = (diff[..., 0] * diff[..., 0] + diff[..., 1] * diff[..., 1]) + diff[..., 2] * diff[..., 2]
    d2s, idx = torch.sort(d2, dim=-1, stable=True)
    dists, vert_ids = d2s[..., :K].sqrt(), idx[..., :K]
    values = values.view(-1, values.shape[-1])
    disp = 1 / (dists + exp)
    weights = disp / disp.sum(dim=-1, keepdim=True)
    dists = torch.einsum('ijk, ijk -> ij', dists, weights)
    sampled = torch.einsum('ijkl, ijk -> ijl', values[vert_ids], weights)
    return sampled.view(n_batch, n_points, -1), dists.view(n_batch, n_points, 1)


# ----------------------------------------------------------------------------
# stage 3: positional encoding  -- embedder.py:5-54
# ----------------------------------------------------------------------------
def positional_encoding(x, n_freq):
    """embedder.py:11-36: [x, sin(2^0 x), cos(2^0 x), ..., sin(2^(L-1) x), cos(2^(L-1) x)]."""
    bands = 2. ** torch.linspace(0., n_freq - 1, steps=n_freq)
    out = [x]
    for f in bands:
        out.append(torch.sin(x * f))
        out.append(torch.cos(x * f))
    return torch.cat(out, -1)


# ----------------------------------------------------------------------------
# stage 2/4: the two MLPs  -- tpose_nerf_network.py
# ----------------------------------------------------------------------------
def _trunk(sd, prefix, feat):
    """8 x (1x1 conv + ReLU), skip concat [input, hidden] after layer 4
    (tpose_nerf_network.py:68-72, 256-260)."""
    net = feat
    for i in range(8):
        net = F.relu(F.conv1d(net, sd[f'{prefix}.{i}.weight'], sd[f'{prefix}.{i}.bias']))
        if i == 4:
            net = torch.cat((feat, net), dim=1)
    return net


def neural_blend_weights(sd, pts, smpl_bw, latent_index, prefix='', xyz_res=10):
    """tpose_nerf_network.py:40-77 (prefix='') and :296-315 (prefix='novel_pose_bw.').
    pts (B,m,3), smpl_bw (B,24,m), latent_index (B,) int64 -> (B,24,m)."""
    x = positional_encoding(pts, xyz_res).transpose(1, 2)
    lat = F.embedding(latent_index, sd[prefix + 'bw_latent.weight'])
    lat = lat[..., None].expand(*lat.shape, x.size(2))
    feat = torch.cat((x, lat), dim=1)
    net = _trunk(sd, prefix + 'bw_linears', feat)
    bw = F.conv1d(net, sd[prefix + 'bw_fc.weight'], sd[prefix + 'bw_fc.bias'])
    bw = torch.log(smpl_bw + 1e-9) + bw
    return F.softmax(bw, dim=1)


def nerf_alpha(sd, pts, xyz_res=10):
    """TPoseHuman.calculate_alpha, tpose_nerf_network.py:241-250."""
    x = positional_encoding(pts, xyz_res).transpose(1, 2)
    net = _trunk(sd, 'tpose_human.pts_linears', x)
    return F.conv1d(net, sd['tpose_human.alpha_fc.weight'], sd['tpose_human.alpha_fc.bias'])


def nerf_alpha_rgb(sd, pts, viewdir, ind, xyz_res=10, view_res=4):
    """TPoseHuman.calculate_alpha_rgb, tpose_nerf_network.py:252-275."""
    p = 'tpose_human.'
    x = positional_encoding(pts, xyz_res).transpose(1, 2)
    net = _trunk(sd, p + 'pts_linears', x)
    alpha = F.conv1d(net, sd[p + 'alpha_fc.weight'], sd[p + 'alpha_fc.bias'])
    feat = F.conv1d(net, sd[p + 'feature_fc.weight'], sd[p + 'feature_fc.bias'])
    lat = F.embedding(ind, sd[p + 'nf_latent.weight'])
    lat = lat[..., None].expand(*lat.shape, net.size(2))
    feat = F.conv1d(torch.cat((feat, lat), dim=1), sd[p + 'latent_fc.weight'], sd[p + 'latent_fc.bias'])
    v = positional_encoding(viewdir, view_res).transpose(1, 2)
    h = F.relu(F.conv1d(torch.cat((feat, v), dim=1), sd[p + 'view_fc.weight'], sd[p + 'view_fc.bias']))
    rgb = F.conv1d(h, sd[p + 'rgb_fc.weight'], sd[p + 'rgb_fc.bias'])
    return alpha, rgb


def pose_to_tpose(sd, pose_pts, batch, cfg):
    """Network.pose_points_to_tpose_points, tpose_nerf_network.py:79-100."""
    init = sample_blend_weights(pose_pts, batch['pbw'], batch['pbounds'])[:, :24]
    if cfg.test_novel_pose:
        pbw = neural_blend_weights(sd, pose_pts, init, batch['bw_latent_index'],
                                   prefix='novel_pose_bw.', xyz_res=cfg.xyz_res)
    else:
        pbw = neural_blend_weights(sd, pose_pts, init, batch['latent_index'] + 1, xyz_res=cfg.xyz_res)
    return inverse_lbs(pose_pts, pbw, batch['A']), pbw


def network_forward(sd, wpts, viewdir, dists, batch, cfg, return_debug=False):
    """Network.forward, tpose_nerf_network.py:139-215.  wpts/viewdir (n,3), dists (n,)."""
    wpts = wpts[None]
    pose_pts = world_to_pose(wpts, batch['R'], batch['Th'])
    init_pbw = sample_blend_weights(pose_pts, batch['pbw'], batch['pbounds'])
    pnorm = init_pbw[:, -1]
    pind = pnorm < cfg.norm_th
    pind[torch.arange(len(pnorm)), pnorm.argmin(dim=1)] = True
    pose_pts = pose_pts[pind][None]
    viewdir = viewdir[pind[0]]
    dists = dists[pind[0]]

    tpose, pbw = pose_to_tpose(sd, pose_pts, batch, cfg)

    init_tbw = sample_blend_weights(tpose, batch['tbw'], batch['tbounds'])[:, :24]
    tbw = neural_blend_weights(sd, tpose, init_tbw, torch.zeros_like(batch['latent_index']),
                               xyz_res=cfg.xyz_res)

    alpha, rgb = nerf_alpha_rgb(sd, tpose, viewdir[None], batch['latent_index'],
                                xyz_res=cfg.xyz_res, view_res=cfg.view_res)

    inside = (tpose > batch['tbounds'][:, :1]) * (tpose < batch['tbounds'][:, 1:])
    outside = torch.sum(inside, dim=2) != 3
    alpha = alpha[:, 0].clone()
    sigma_raw = alpha.clone()
    alpha[outside] = 0

    alpha_ind = alpha > cfg.train_th
    alpha_ind[torch.arange(alpha.size(0)), torch.argmax(alpha, dim=1)] = True
    pbw_sel = pbw.transpose(1, 2)[alpha_ind][None]
    tbw_sel = tbw.transpose(1, 2)[alpha_ind][None]

    rgb_s = torch.sigmoid(rgb[0])
    a = 1. - torch.exp(-F.relu(alpha[0]) * dists)
    raw = torch.cat((rgb_s, a[None]), dim=0).transpose(0, 1)
    raw_full = torch.zeros([1, wpts.shape[1], 4], dtype=wpts.dtype, device=wpts.device)
    raw_full[pind] = raw
    ret = {'pbw': pbw_sel, 'tbw': tbw_sel, 'raw': raw_full}
    if return_debug == 'light':
        # full-size frames: only what the row-selection / active-set checks need (no per-sample (n,25) tensors)
        ret['_debug'] = {'pind': pind[0], 'sigma_masked': alpha[0], 'alpha_ind': alpha_ind[0], 'tpose': tpose[0]}
    elif return_debug:
        ret['_debug'] = {'pind': pind[0], 'pnorm': pnorm[0], 'pose_pts': pose_pts[0], 'tpose': tpose[0],
                         'pbw_all': pbw[0].transpose(0, 1), 'tbw_all': tbw[0].transpose(0, 1),
                         'init_pbw': init_pbw[0].transpose(0, 1), 'sigma': sigma_raw[0], 'rgb_raw': rgb[0].transpose(0, 1),
                         'outside': outside[0], 'alpha_ind': alpha_ind[0], 'sigma_masked': alpha[0]}
    return ret


def calculate_alpha(sd, wpts, batch, cfg):
    """Network.calculate_alpha (= get_alpha), tpose_nerf_network.py:105-137.  norm_th is 0.1 here."""
    wpts = wpts[None]
    pose_pts = world_to_pose(wpts, batch['R'], batch['Th'])
    pnorm = sample_blend_weights(pose_pts, batch['pbw'], batch['pbounds'])[:, 24]
    pind = pnorm < 0.1
    pind[torch.arange(len(pnorm)), pnorm.argmin(dim=1)] = True
    pose_pts = pose_pts[pind][None]
    tpose, _ = pose_to_tpose(sd, pose_pts, batch, cfg)
    alpha = nerf_alpha(sd, tpose, xyz_res=cfg.xyz_res)[0, 0]
    full = torch.zeros([wpts.shape[1]]).to(wpts)
    full[pind[0]] = alpha
    return full


# ----------------------------------------------------------------------------
# stage 5: compositing  -- nerf_net_utils.py:6-36
# ----------------------------------------------------------------------------
def raw2outputs(raw, z_vals, white_bkgd=False):
    rgb = raw[..., :-1]
    alpha = raw[..., -1]
    T = torch.cumprod(torch.cat([torch.ones((alpha.shape[0], 1)).to(alpha), 1. - alpha + 1e-10], -1), -1)[:, :-1]
    w = alpha * T
    rgb_map = torch.sum(w[..., None] * rgb, -2)
    depth_map = torch.sum(w * z_vals, -1)
    disp_map = 1. / torch.max(1e-10 * torch.ones_like(depth_map), depth_map / torch.sum(w, -1))
    acc_map = torch.sum(w, -1)
    if white_bkgd:
        rgb_map = rgb_map + (1. - acc_map[..., None])
    return rgb_map, disp_map, acc_map, w, depth_map


# ----------------------------------------------------------------------------
# the renderer  -- tpose_renderer.py:71-186
# ----------------------------------------------------------------------------
def render(sd, batch, cfg=None, t_rand=None, return_debug=False, grad=False):
    """Renderer.render: 2048-ray chunks -> sample -> Network.forward -> raw2outputs; concat on dim 1.
    grad=True keeps the autograd graph (training, tpose_renderer.py:154 keeps device tensors with grad)."""
    cfg = cfg or OracleCfg()
    ray_o, ray_d, near, far = batch['ray_o'], batch['ray_d'], batch['near'], batch['far']
    R = ray_o.shape[1]
    S = cfg.N_samples
    outs = []
    with torch.set_grad_enabled(bool(grad)):
        for i in range(0, R, cfg.chunk):
            o, d = ray_o[:, i:i + cfg.chunk], ray_d[:, i:i + cfg.chunk]
            tr = None if t_rand is None else t_rand[:, i:i + cfg.chunk]
            pts, z = sample_points(o, d, near[:, i:i + cfg.chunk], far[:, i:i + cfg.chunk], S, tr)
            nb, npix = pts.shape[:2]
            view = d[:, :, None].repeat(1, 1, S, 1).contiguous().view(nb * npix * S, -1)
            ret = network_forward(sd, pts.view(nb * npix * S, -1), view, sample_dists(z).view(-1),
                                  batch, cfg, return_debug=return_debug)
            raw = ret['raw'].reshape(-1, S, 4)
            rgb_map, disp, acc, w, depth = raw2outputs(raw, z.view(-1, S), cfg.white_bkgd)
            o_ = {'rgb_map': rgb_map.view(nb, npix, -1), 'acc_map': acc.view(nb, npix),
                  'depth_map': depth.view(nb, npix), 'raw': raw.view(nb, -1, 4),
                  'pbw': ret['pbw'].view(nb, -1, 24), 'tbw': ret['tbw'].view(nb, -1, 24)}
            if return_debug:
                o_['_debug'] = ret['_debug']
                if return_debug != 'light':
                    o_['_debug']['z_vals'] = z[0]
                    o_['_debug']['wpts'] = pts[0]
            outs.append(o_)
    out = {k: torch.cat([r[k] for r in outs], dim=1) for k in outs[0] if k != '_debug'}
    if return_debug:
        dbg = {}
        for k in outs[0]['_debug']:
            dbg[k] = torch.cat([r['_debug'][k] for r in outs], dim=0)
        dbg['chunk_active'] = torch.tensor([int(r['_debug']['pind'].sum()) for r in outs])
        out['_debug'] = dbg
    return out


# ----------------------------------------------------------------------------
# the training step  -- lib/train/trainers/tpose_trainer.py:21-73, trainer.py:54-68
# ----------------------------------------------------------------------------
def train_loss(sd, batch, cfg=None, t_rand=None):
    """NetworkWrapper.forward: loss = smooth_l1(pbw, tbw) + mse(rgb_map[mask], rgb[mask]).  `sd` holds the
    parameters (requires_grad leaves for gradients).  Returns (ret, loss, scalar_stats)."""
    cfg = cfg or OracleCfg()
    ret = render(sd, batch, cfg, t_rand=t_rand, grad=True)
    bw_loss = F.smooth_l1_loss(ret['pbw'], ret['tbw'])
    mask = batch['mask_at_box']
    img_loss = torch.mean((ret['rgb_map'][mask] - batch['rgb'][mask]) ** 2)
    loss = bw_loss + img_loss
    return ret, loss, {'bw_loss': bw_loss, 'img_loss': img_loss, 'loss': loss}


def train_step_grads(sd, batch, cfg=None, t_rand=None, clip=40.0):
    """One Trainer.train iteration up to the optimizer (trainer.py:62-66): zero_grad, backward,
    clip_grad_value_(40).  Returns (loss stats as floats, {name: grad})."""
    params = {k: v.detach().clone().requires_grad_(True) for k, v in sd.items()}
    _, loss, stats = train_loss(params, batch, cfg, t_rand)
    loss.backward()
    grads = {k: (v.grad.clamp(-clip, clip) if v.grad is not None else torch.zeros_like(v)) for k, v in params.items()}
    return {k: float(v) for k, v in stats.items()}, grads


# ----------------------------------------------------------------------------
# stage 2: novel-pose blend-weight training  -- lib/train/trainers/aninerf_animation_trainer.py:11-152
# ----------------------------------------------------------------------------
def _select_rows(alpha, train_th):
    ind = alpha.detach() > train_th
    ind[torch.arange(alpha.size(0)), torch.argmax(alpha, dim=1)] = True
    return ind


def animation_train_loss(sd, batch, wpts, tpts, cfg=None):
    """NetworkWrapper.forward of the second training stage.  wpts / tpts (1,n,3): the uniform samples of wbounds / tbounds
    that get_sampling_points draws with torch.rand (:122-142), passed in for parity.  Only `novel_pose_bw.*` is trained."""
    cfg = cfg or OracleCfg()
    idx = batch['bw_latent_index']
    zero = torch.zeros_like(idx)
    # observation space -> canonical (ppts_to_tpose, :56-91)
    ppts = world_to_pose(wpts, batch['R'], batch['Th'])
    pv = sample_blend_weights(ppts, batch['pbw'], batch['pbounds'])
    init_pbw, pnorm = pv[:, :24], pv[:, 24]
    pbw = neural_blend_weights(sd, ppts, init_pbw, idx, prefix='novel_pose_bw.', xyz_res=cfg.xyz_res)
    tpose = inverse_lbs(ppts, pbw, batch['A'])
    init_tbw = sample_blend_weights(tpose, batch['tbw'], batch['tbounds'])[:, :24]
    tbw = neural_blend_weights(sd, tpose, init_tbw, zero, xyz_res=cfg.xyz_res)
    alpha = nerf_alpha(sd, tpose, xyz_res=cfg.xyz_res)
    inside = torch.sum((tpose > batch['tbounds'][:, :1]) * (tpose < batch['tbounds'][:, 1:]), dim=2) == 3
    inside = inside * (pnorm < cfg.norm_th)
    alpha = alpha[:, 0].clone()
    alpha[~inside] = 0
    sel0 = _select_rows(alpha, cfg.train_th)
    pbw0, tbw0 = pbw.transpose(1, 2)[sel0], tbw.transpose(1, 2)[sel0]
    # canonical space -> observation (tpose_to_ppts, :94-119)
    init_tbw1 = sample_blend_weights(tpts, batch['tbw'], batch['tbounds'])[:, :24]
    tbw_c = neural_blend_weights(sd, tpts, init_tbw1, zero, xyz_res=cfg.xyz_res)
    alpha1 = nerf_alpha(sd, tpts, xyz_res=cfg.xyz_res)[:, 0]
    pose_pts = forward_lbs(tpts, tbw_c, batch['A'])
    init_pbw1 = sample_blend_weights(pose_pts, batch['pbw'], batch['pbounds'])[:, :24]
    pbw_c = neural_blend_weights(sd, pose_pts, init_pbw1, idx, prefix='novel_pose_bw.', xyz_res=cfg.xyz_res)
    sel1 = _select_rows(alpha1, cfg.train_th)
    pbw1, tbw1 = pbw_c.transpose(1, 2)[sel1], tbw_c.transpose(1, 2)[sel1]
    l0 = F.smooth_l1_loss(pbw0, tbw0)
    l1 = F.smooth_l1_loss(pbw1, tbw1)
    loss = l0 + l1
    return {'pbw0': pbw0}, loss, {'bw_loss0': l0, 'bw_loss1': l1, 'loss': loss}


def animation_train_step_grads(sd, batch, wpts, tpts, cfg=None, clip=40.0):
    """One stage-2 iteration up to the optimizer: only novel_pose_bw.* requires grad (:26-31)."""
    params = {k: (v.detach().clone().requires_grad_(k.startswith('novel_pose_bw.'))) for k, v in sd.items()}
    _, loss, stats = animation_train_loss(params, batch, wpts, tpts, cfg)
    loss.backward()
    grads = {k: v.grad.clamp(-clip, clip) for k, v in params.items() if k.startswith('novel_pose_bw.') and v.grad is not None}
    return {k: float(v) for k, v in stats.items()}, grads


# ----------------------------------------------------------------------------
# novel-view / pose-sequence renderer with multi-view silhouette culling
# -- lib/networks/renderer/tpose_renderer_mmsk.py:14-166
# ----------------------------------------------------------------------------
def inside_all_views(pts, batch):
    """prepare_inside_pts, tpose_renderer_mmsk.py:14-57.  pts (1,m,3) -> bool (1,m): the sample projects
    into the (dilated) silhouette of EVERY training view."""
    H, W = int(batch['H']), int(batch['W'])
    inside = None
    for nv in range(batch['Ks'].size(1)):
        R = batch['RT'][:, nv, :3, :3]
        T = batch['RT'][:, nv, :3, 3]
        p = torch.matmul(pts, R.transpose(2, 1)) + T[:, None]
        p = torch.matmul(p, batch['Ks'][:, nv].transpose(2, 1))
        xy = (p[..., :2] / p[..., 2:]).round().long()
        xy[..., 0] = torch.clamp(xy[..., 0], 0, W - 1)
        xy[..., 1] = torch.clamp(xy[..., 1], 0, H - 1)
        xy = xy[0]
        m = batch['msks'][0, nv][xy[:, 1], xy[:, 0]][None].bool()
        inside = m if inside is None else inside * m
    return inside


def render_mmsk(sd, batch, cfg=None, return_debug=False):
    """tpose_renderer_mmsk.Renderer.render: per 2048-ray chunk, cull samples outside any training-view
    silhouette, run Network.forward on the survivors only, composite.  Returns rgb/acc/depth maps."""
    cfg = cfg or OracleCfg()
    ray_o, ray_d, near, far = batch['ray_o'], batch['ray_d'], batch['near'], batch['far']
    R = ray_o.shape[1]
    S = cfg.N_samples
    outs, dbg_inside, dbg_active = [], [], []
    with torch.no_grad():
        for i in range(0, R, cfg.chunk):
            o, d = ray_o[:, i:i + cfg.chunk], ray_d[:, i:i + cfg.chunk]
            pts, z = sample_points(o, d, near[:, i:i + cfg.chunk], far[:, i:i + cfg.chunk], S)
            nb, npix = pts.shape[:2]
            inside = inside_all_views(pts.view(nb, -1, 3), batch).view(-1)
            full_raw = torch.zeros([nb * npix * S, 4], device=pts.device)
            n_act = 0
            if inside.sum() > 0:
                w = pts.view(-1, 3)[inside]
                view = d[:, :, None].repeat(1, 1, S, 1).contiguous().view(-1, 3)[inside]
                dist = sample_dists(z).view(-1)[inside]
                ret = network_forward(sd, w, view, dist, batch, cfg, return_debug=True)
                full_raw[inside] = ret['raw'][0]
                n_act = int(ret['_debug']['pind'].sum())
            raw = full_raw.reshape(-1, S, 4)
            rgb_map, disp, acc, wgt, depth = raw2outputs(raw, z.view(-1, S), cfg.white_bkgd)
            outs.append({'rgb_map': rgb_map.view(nb, npix, -1), 'acc_map': acc.view(nb, npix), 'depth_map': depth.view(nb, npix)})
            dbg_inside.append(inside)
            dbg_active.append(n_act)
    out = {k: torch.cat([r[k] for r in outs], dim=1) for k in outs[0]}
    if return_debug:
        out['_debug'] = {'inside': torch.cat(dbg_inside), 'chunk_active': torch.tensor(dbg_active)}
    return out
