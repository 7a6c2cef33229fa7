"""On-disk formats at the edges of the path: the reference's `.pth` checkpoints and the per-frame `.npy`
files of a processed ZJU-MoCap / H36M sequence.

 * `save_model / load_model / load_network` write and read the reference's checkpoint dict `{'net', 'optim', 'scheduler',
   'recorder', 'epoch'}` under `<model_dir>/<epoch>.pth` / `latest.pth` (lib/utils/net_utils.py:288-396), so a run can be
   resumed by either code base; `Network` (tpose_nerf_network.py) keeps the reference's parameter names, which is all the
   format asks of it.
 * `load_frame` assembles the frame part of a render batch from the files `lib/datasets/tpose_dataset.py:125-161, 205-222`
   reads: `<data_root>/<vertices>/<i>.npy`, `<data_root>/<params>/<i>.npy` (dict: Rh, Th, poses, shapes),
   `<lbs_root>/bweights/<i>.npy`, `tbw.npy`, `tvertices.npy`, `joints.npy`, `parents.npy`.
"""
from __future__ import annotations

import os
import shutil

import numpy as np
import torch

from . import host_geometry

KEEP_LAST = 20      # net_utils.py:341-349: only the newest 20 numbered checkpoints are kept


def _numbered(model_dir):
    return sorted(int(f.split('.')[0]) for f in os.listdir(model_dir) if f.endswith('.pth') and f != 'latest.pth' and f.split('.')[0].isdigit())


def _pick(model_dir, epoch):
    """path of the checkpoint `epoch` (-1: latest.pth if present, else the highest number) or None"""
    if not os.path.isdir(model_dir):
        return model_dir if os.path.isfile(model_dir) else None
    nums = _numbered(model_dir)
    has_latest = os.path.exists(os.path.join(model_dir, 'latest.pth'))
    if epoch == -1:
        if has_latest:
            return os.path.join(model_dir, 'latest.pth')
        return os.path.join(model_dir, f'{nums[-1]}.pth') if nums else None
    return os.path.join(model_dir, f'{epoch}.pth')


def save_model(net, optim, scheduler, recorder, model_dir, epoch, last=False):
    os.makedirs(model_dir, exist_ok=True)
    state = {'net': net.state_dict(), 'optim': optim.state_dict(), 'scheduler': scheduler.state_dict() if scheduler is not None else {},
             'recorder': recorder.state_dict() if recorder is not None else {}, 'epoch': epoch}
    torch.save(state, os.path.join(model_dir, 'latest.pth' if last else f'{epoch}.pth'))
    nums = _numbered(model_dir)
    if len(nums) > KEEP_LAST:
        os.remove(os.path.join(model_dir, f'{nums[0]}.pth'))


def load_model(net, optim, scheduler, recorder, model_dir, resume=True, epoch=-1):
    """-> the epoch to continue with (0 when there is nothing to resume).  `resume=False` removes `model_dir` first, as the
    reference does (`os.system('rm -rf ...')`, lib/utils/net_utils.py:295-296): a fresh run must not inherit the previous run's
    `latest.pth` or have its numbered checkpoints pruned together with the old ones."""
    if not resume:
        shutil.rmtree(model_dir, ignore_errors=True)
        return 0
    if not os.path.exists(model_dir):
        return 0
    path = _pick(model_dir, epoch)
    if path is None:
        return 0
    state = torch.load(path, map_location='cpu')
    net.load_state_dict(state['net'])
    optim.load_state_dict(state['optim'])
    if scheduler is not None and state.get('scheduler'):
        scheduler.load_state_dict(state['scheduler'])
    if recorder is not None and state.get('recorder'):
        recorder.load_state_dict(state['recorder'])
    return state['epoch'] + 1


def load_network(net, model_dir, resume=True, epoch=-1, strict=True, only=()):
    """Weights only; `only` = key prefixes to keep (implies strict=False), as net_utils.load_network."""
    if not resume or not os.path.exists(model_dir):
        return 0
    path = _pick(model_dir, epoch)
    if path is None:
        return 0
    state = torch.load(path, map_location='cpu')
    sd = state['net']
    if only:
        strict = False
        sd = {k: v for k, v in sd.items() if any(k.startswith(p) for p in only)}
    net.load_state_dict(sd, strict=strict)
    return state['epoch'] + 1


def load_frame(data_root, lbs_root, i, latent_index=0, bw_latent_index=0, vertices='new_vertices', params='new_params', box_padding=0.05):
    """Frame dict (numpy, no batch dim) with the keys of `synthetic.FRAME_KEYS` from a processed sequence on disk."""
    wxyz = np.load(os.path.join(data_root, vertices, f'{i}.npy')).astype(np.float32)
    prm = np.load(os.path.join(data_root, params, f'{i}.npy'), allow_pickle=True).item()
    Rh = np.asarray(prm['Rh'], dtype=np.float32).reshape(3)
    Th = np.asarray(prm['Th'], dtype=np.float32).reshape(1, 3)
    R = host_geometry.axis_angle_to_matrix(Rh).astype(np.float32)
    pxyz = np.dot(wxyz - Th, R).astype(np.float32)
    joints = np.load(os.path.join(lbs_root, 'joints.npy')).astype(np.float32)
    parents = np.load(os.path.join(lbs_root, 'parents.npy'))
    A = host_geometry.bone_transforms(np.asarray(prm['poses']).reshape(-1, 3), joints, parents)
    tverts = np.load(os.path.join(lbs_root, 'tvertices.npy')).astype(np.float32)
    return {
        'A': A.astype(np.float32), 'R': R, 'Th': Th,
        'pbw': np.load(os.path.join(lbs_root, 'bweights', f'{i}.npy')).astype(np.float32),
        'tbw': np.load(os.path.join(lbs_root, 'tbw.npy')).astype(np.float32),
        'pbounds': host_geometry.bounds_of(pxyz, box_padding), 'wbounds': host_geometry.bounds_of(wxyz, box_padding),
        'tbounds': host_geometry.bounds_of(tverts, box_padding),
        'latent_index': np.int64(latent_index), 'bw_latent_index': np.int64(bw_latent_index),
    }
