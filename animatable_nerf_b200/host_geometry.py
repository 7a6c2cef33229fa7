"""Host-side (numpy, float64) inputs of the render path.

These produce the *inputs* of the hot path (SMPL bone transforms A, bounds, rigid
frame R/Th, camera extrinsics); they are tiny per-frame computations that stay on
the host, as in the reference's Dataset (`lib/utils/if_nerf/if_nerf_data_utils.py`
:392-458 `batch_rodrigues` / `get_rigid_transformation`, :566-579 `get_bounds`;
`lib/datasets/tpose_dataset.py`:148 `cv2.Rodrigues`).  Per-ray work (ray
generation, box intersection) is NOT here: that runs on the GPU
(`csrc/geometry.cu`: `gen_rays_kernel`, `near_far_kernel`; host mirror `frontend.py`).
"""
from __future__ import annotations

import numpy as np

# SMPL kinematic tree (24 joints); parents[0] is never dereferenced.
SMPL_PARENTS = np.array([-1, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17, 18, 19, 20, 21])


def bounds_of(xyz: np.ndarray, box_padding: float = 0.05) -> np.ndarray:
    """(V,3) -> (2,3) float32 axis-aligned box grown by `box_padding` (cfg.box_padding)."""
    xyz = np.asarray(xyz)
    lo = xyz.min(axis=0) - box_padding
    hi = xyz.max(axis=0) + box_padding
    return np.stack([lo, hi]).astype(np.float32)


def axis_angle_to_matrix(rvec: np.ndarray) -> np.ndarray:
    """Rodrigues formula for one axis-angle vector -> (3,3) float64 (the role cv2.Rodrigues
    plays at tpose_dataset.py:148)."""
    r = np.asarray(rvec, dtype=np.float64).reshape(3)
    th = np.linalg.norm(r)
    if th < 1e-12:
        return np.eye(3)
    k = r / th
    Kx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + np.sin(th) * Kx + (1 - np.cos(th)) * (Kx @ Kx)


def axis_angles_to_matrices(poses: np.ndarray) -> np.ndarray:
    """(J,3) -> (J,3,3); same regularisation as the reference (norm of poses+1e-8)."""
    poses = np.asarray(poses, dtype=np.float64)
    th = np.linalg.norm(poses + 1e-8, axis=1, keepdims=True)
    k = poses / th
    c, s = np.cos(th)[:, None], np.sin(th)[:, None]
    Z = np.zeros(len(poses))
    Kx = np.stack([Z, -k[:, 2], k[:, 1], k[:, 2], Z, -k[:, 0], -k[:, 1], k[:, 0], Z], axis=1).reshape(-1, 3, 3)
    return np.eye(3)[None] + s * Kx + (1 - c) * (Kx @ Kx)


def bone_transforms(poses: np.ndarray, joints: np.ndarray, parents: np.ndarray = SMPL_PARENTS) -> np.ndarray:
    """A_k = G_k(pose) G_k(rest)^-1 for the 24 SMPL bones -> (24,4,4) float32.

    Walk the kinematic chain with local transforms [R_k | j_k - j_parent(k)], then remove the
    rest pose by subtracting G_k[:3,:3] j_k from the translation column.
    """
    rot = axis_angles_to_matrices(poses)
    joints = np.asarray(joints, dtype=np.float64)
    G = np.zeros((len(parents), 4, 4))
    for k in range(len(parents)):
        local = np.eye(4)
        local[:3, :3] = rot[k]
        local[:3, 3] = joints[k] - (joints[parents[k]] if k > 0 else 0.0)
        G[k] = local if k == 0 else G[parents[k]] @ local
    A = G.copy()
    A[:, :3, 3] -= np.einsum('kij,kj->ki', G[:, :3, :3], joints)
    return A.astype(np.float32)


def look_at_camera(center: np.ndarray, distance: float, azimuth: float = 0.0, up=(0.0, 0.0, 1.0)):
    """World->camera (R,T) of a pinhole camera on a circle of radius `distance` around `center`,
    looking at it (OpenCV convention: +z forward, +y down).  Used to synthesise test cameras; a
    64-view sweep is `azimuth = 2*pi*i/64` (cf. lib/utils/render_utils.py:75-127 `gen_path`)."""
    center = np.asarray(center, dtype=np.float64)
    up = np.asarray(up, dtype=np.float64)
    a = np.array([np.cos(azimuth), np.sin(azimuth), 0.0])
    b = np.cross(up, a)
    pos = center + distance * (a * 1.0 + b * 0.0)
    fwd = center - pos
    fwd /= np.linalg.norm(fwd)
    right = np.cross(fwd, up)
    right /= np.linalg.norm(right)
    down = np.cross(fwd, right)
    R = np.stack([right, down, fwd])            # rows = camera axes in world coords
    T = -R @ pos
    return R, T.reshape(3, 1)


# ---------------------------------------------------------------------------------------------
# novel-view camera path (lib/utils/render_utils.py:11-28, 77-127 `gen_path`)
# ---------------------------------------------------------------------------------------------
def _unit(v):
    return v / np.linalg.norm(v)


def circular_camera_path(RT, render_views: int, center=None):
    """World->camera matrices of a `render_views`-step circular sweep around the capture rig, from the rig's
    world->camera matrices `RT` (c,4,4): the path render_utils.gen_path builds for `tpose_novel_view_dataset`.
    Camera axes follow the LLFF convention used there ([down, right, backwards]); the path is an ellipse whose radii
    are 1.3 x the 80th percentile of the rig cameras' offsets in the mean camera frame.  Host, float64, once per sweep."""
    c2w_all = np.linalg.inv(np.array(RT, dtype=np.float64))
    c2w_all = np.concatenate([c2w_all[:, :, 1:2], c2w_all[:, :, 0:1], -c2w_all[:, :, 2:3], c2w_all[:, :, 3:4]], 2)
    up = _unit(c2w_all[:, :3, 0].sum(0))
    gaze = _unit(c2w_all[0, :3, 2])
    side = _unit(np.cross(gaze, up))
    fwd = _unit(np.cross(up, side))
    z_off = 0
    if center is None:
        center = c2w_all[:, :3, 3].mean(0)
        z_off = 1.3
    frame = np.stack([up, side, fwd, center], 1)                                   # (3,4) mean camera -> world
    local = np.matmul(frame[:3, :3].T, (c2w_all[:, :3, 3] - frame[:3, 3])[..., np.newaxis])[..., 0].T
    rads = np.percentile(np.abs(local), 80, -1) * 1.3
    rads = np.array(list(rads) + [1.])
    target = np.dot(frame[:3, :4], np.array([z_off, 0, 0, 1.]))
    out = []
    for theta in np.linspace(0., 2 * np.pi, render_views + 1)[:-1]:
        pos = np.dot(frame[:3, :4], np.array([0, np.sin(theta), np.cos(theta), 1] * rads))
        z = _unit(pos - target)
        v1 = _unit(np.cross(z, up))
        v0 = _unit(np.cross(v1, z))
        m = np.stack([v0, v1, z, pos], 1)
        m = np.concatenate([m[:, 1:2], m[:, 0:1], -m[:, 2:3], m[:, 3:4]], 1)
        m = np.concatenate([m, np.array([[0., 0., 0., 1.]])], 0)
        out.append(np.linalg.inv(m))
    return out
