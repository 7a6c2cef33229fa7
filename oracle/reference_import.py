"""TEST INFRASTRUCTURE ONLY -- import the UNMODIFIED reference from /root/reference on CPU.

Only usable in the build container (the GPU box has no /root/reference).  Used by
`oracle/validate_against_reference.py` to pin the oracle restatement and to generate the
golden fixtures under tests/golden/.  Recipe: SURVEY.md section 8c (stub the modules the
path imports but never calls, set sys.argv before `lib.config` is imported, chdir).
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

REF = '/root/reference'


def available() -> bool:
    return os.path.isdir(os.path.join(REF, 'lib'))


def _stub(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def _load_source(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


_loaded = None


def load(cfg_file='configs/aninerf_313.yaml', overrides=()):
    """Returns a namespace with cfg, Network, Renderer, data utils and nerf_net_utils of the reference."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError('/root/reference is not present on this machine')
    saved_argv, saved_cwd = list(sys.argv), os.getcwd()
    saved_cvd = os.environ.get('CUDA_VISIBLE_DEVICES')
    _stub('imp', load_source=_load_source)
    _stub('termcolor', colored=lambda s, *a, **k: s)
    _stub('pytorch3d', _C=None)
    _stub('pytorch3d.structures', Meshes=None, Pointclouds=None)
    _stub('pytorch3d.ops', knn_points=None, knn_gather=None, sample_points_from_meshes=None)
    _stub('pytorch3d.ops.packed_to_padded', packed_to_padded=None)
    _stub('pytorch3d.ops.knn', knn_gather=None, knn_points=None)
    _stub('pytorch3d.ops.mesh_face_areas_normals', mesh_face_areas_normals=None)
    _stub('pytorch3d.ops.sample_points_from_meshes', sample_points_from_meshes=None,
          _rand_barycentric_coords=None)
    for name in ('trimesh', 'imageio', 'tensorboardX'):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                _stub(name, SummaryWriter=None)
    os.chdir(REF)
    sys.path.insert(0, REF)
    sys.argv = ['oracle', '--cfg_file', cfg_file, 'exp_name', 'oracle', 'resume', 'False'] + list(overrides)
    try:
        from lib.config import cfg
        net_mod = _load_source(cfg.network_module, cfg.network_path)
        ren_mod = _load_source(cfg.renderer_module, cfg.renderer_path)
        from lib.utils.if_nerf import if_nerf_data_utils as dutils
        from lib.networks.renderer import nerf_net_utils
        from lib.utils import blend_utils
        from lib.networks import embedder
        from lib.networks.renderer import tpose_renderer_mmsk as mmsk_mod
        render_utils = None
        try:
            from lib.utils import render_utils
        except Exception as e:  # noqa: BLE001
            print('reference_import: render_utils not importable:', e)
        anim_trainer_mod = None
        try:
            from lib.train.trainers import aninerf_animation_trainer as anim_trainer_mod
        except Exception as e:  # noqa: BLE001
            print('reference_import: aninerf_animation_trainer not importable:', e)
        trainer_mod = None
        try:
            from lib.train.trainers import tpose_trainer as trainer_mod
        except Exception as e:  # noqa: BLE001  (optional: needs the tensorboardX / trimesh stubs)
            print('reference_import: tpose_trainer not importable:', e)
    finally:
        sys.argv = saved_argv
        os.chdir(saved_cwd)
        if saved_cvd is None:
            os.environ.pop('CUDA_VISIBLE_DEVICES', None)
        else:
            os.environ['CUDA_VISIBLE_DEVICES'] = saved_cvd
    _loaded = types.SimpleNamespace(cfg=cfg, Network=net_mod.Network, Renderer=ren_mod.Renderer,
                                    net_mod=net_mod, dutils=dutils, nerf_net_utils=nerf_net_utils,
                                    blend_utils=blend_utils, embedder=embedder, MmskRenderer=mmsk_mod.Renderer, trainer_mod=trainer_mod, render_utils=render_utils, anim_trainer_mod=anim_trainer_mod)
    return _loaded
