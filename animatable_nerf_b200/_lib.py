"""ctypes binding of libaninerf_b200.so (the C ABI declared in include/aninerf_b200.h).

There is no CPU fallback: if the shared library is missing or a call fails, this raises.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libaninerf_b200.so')

N_BONES = 24
BW_CH = 25
CHUNK_RAYS = 2048
FIELD_BW, FIELD_NOVEL_BW, FIELD_NERF = 0, 1, 2
STAGES = ('split_volumes', 'clear_raw', 'front_end', 'unused3', 'unused4', 'bw_field_posed', 'bw_field_canonical', 'nerf_field',
          'composite')

c_float_p = C.POINTER(C.c_float)
c_i32_p = C.POINTER(C.c_int32)
c_u8_p = C.POINTER(C.c_uint8)


class Camera(C.Structure):
    _fields_ = [('Kinv', C.c_double * 9), ('R', C.c_double * 9), ('T', C.c_double * 3), ('H', C.c_int32), ('W', C.c_int32)]


class Layer(C.Structure):
    _fields_ = [('W', c_float_p), ('bias_table', c_float_p), ('n_out', C.c_int32), ('k_in', C.c_int32),
                ('n_tables', C.c_int32), ('relu', C.c_int32)]


class Frame(C.Structure):
    _fields_ = [('A', C.c_void_p), ('R', C.c_void_p), ('Th', C.c_void_p), ('pbw', C.c_void_p), ('tbw', C.c_void_p),
                ('pbounds', C.c_void_p), ('tbounds', C.c_void_p), ('pbw_dims', C.c_int32 * 3), ('tbw_dims', C.c_int32 * 3),
                ('latent_index', C.c_int32), ('bw_latent_index', C.c_int32), ('latent_index_dev', C.c_void_p),
                ('bw_latent_index_dev', C.c_void_p)]


class RenderParams(C.Structure):
    _fields_ = [('n_samples', C.c_int32), ('chunk_rays', C.c_int32), ('norm_th', C.c_float), ('white_bkgd', C.c_int32),
                ('novel_pose', C.c_int32), ('want_bw', C.c_int32), ('bw_precision', C.c_int32), ('nerf_precision', C.c_int32)]


class RenderOutputs(C.Structure):
    _fields_ = [('rgb_map', C.c_void_p), ('acc_map', C.c_void_p), ('depth_map', C.c_void_p), ('raw', C.c_void_p),
                ('pbw_all', C.c_void_p), ('tbw_all', C.c_void_p), ('sigma_masked', C.c_void_p), ('active_index', C.c_void_p),
                ('n_active', C.c_void_p), ('chunk_offsets', C.c_void_p)]


class PeerGather(C.Structure):
    _fields_ = [('maps', C.c_void_p * 8), ('world', C.c_int32), ('rank', C.c_int32)]


class Silhouettes(C.Structure):
    _fields_ = [('msks', C.c_void_p), ('Ks', C.c_void_p), ('RT', C.c_void_p), ('n_views', C.c_int32), ('H', C.c_int32), ('W', C.c_int32)]


class GemmSeg(C.Structure):
    _fields_ = [('A', C.c_void_p), ('a_row_stride', C.c_int64), ('a_k_stride', C.c_int64), ('B', C.c_void_p), ('b_row_stride', C.c_int64),
                ('b_k_stride', C.c_int64), ('K', C.c_int32)]


class Gemm(C.Structure):
    _fields_ = [('seg', GemmSeg * 2), ('n_seg', C.c_int32), ('M', C.c_int32), ('N', C.c_int32), ('C', C.c_void_p), ('ldc', C.c_int64),
                ('bias', C.c_void_p), ('relu_mask', C.c_void_p), ('ld_mask', C.c_int64), ('relu', C.c_int32), ('accumulate', C.c_int32),
                ('split_k', C.c_int32)]


# name -> (restype, argtypes); every symbol include/aninerf_b200.h declares
_VP, _I32, _I64, _F = C.c_void_p, C.c_int32, C.c_int64, C.c_float
PROTOTYPES = {
    'aninerf_version': (_I32, []),
    'aninerf_last_error': (C.c_char_p, []),
    'aninerf_launch_count': (_I64, []),
    'aninerf_debug_set_trace': (_I32, [_VP]),
    'aninerf_profile_enable': (_I32, [_I32]),
    'aninerf_profile_read': (_I32, [C.POINTER(C.c_double), C.POINTER(C.c_int64), _I32]),
    'aninerf_gen_rays': (_I32, [C.POINTER(Camera), _VP, _VP, _VP]),
    'aninerf_near_far': (_I32, [c_float_p, _VP, _VP, _I64, _VP, _VP, _VP, _VP]),
    'aninerf_compact_workspace_bytes': (_I64, [_I64]),
    'aninerf_compact_rays': (_I32, [_VP, _VP, _VP, _VP, _VP, _I64, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _I64, _VP]),
    'aninerf_sample_points': (_I32, [_VP, _VP, _VP, _VP, _VP, _VP, _I64, _I32, _VP, _VP, _VP, _VP]),
    'aninerf_world_to_pose': (_I32, [_VP, _I64, _VP, _VP, _VP, _VP]),
    'aninerf_sample_blend_weights': (_I32, [_VP, _I64, _VP, C.POINTER(_I32), _VP, _VP, _VP]),
    'aninerf_inverse_lbs': (_I32, [_VP, _VP, _I64, _VP, _VP, _VP]),
    'aninerf_forward_lbs': (_I32, [_VP, _VP, _I64, _VP, _VP, _VP]),
    'aninerf_composite': (_I32, [_VP, _VP, _I64, _I32, _I32, _VP, _VP, _VP, _VP, _VP, _VP]),
    'aninerf_net_create': (_I32, [C.POINTER(_VP)]),
    'aninerf_net_destroy': (_I32, [_VP]),
    'aninerf_net_load_field': (_I32, [_VP, _I32, C.POINTER(Layer), _I32, c_float_p, c_float_p, c_float_p, c_float_p, _VP]),
    'aninerf_bw_forward': (_I32, [_VP, _I32, _I32, _VP, _VP, _I64, _VP, _VP, _VP, _VP, _I32, _VP]),
    'aninerf_nerf_forward': (_I32, [_VP, _I32, _VP, _VP, _I64, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _I32, _VP]),
    'aninerf_render_workspace_bytes': (_I64, [_I64, _I32, _I32, _I64, _I64]),
    'aninerf_render_rays': (_I32, [_VP, C.POINTER(Frame), C.POINTER(RenderParams), _VP, _VP, _VP, _VP, _VP, _VP, _I64,
                                   C.POINTER(RenderOutputs), _VP, _I64, _VP]),
    'aninerf_render_rays_culled': (_I32, [_VP, C.POINTER(Frame), C.POINTER(RenderParams), C.POINTER(Silhouettes), _VP, _VP, _VP, _VP, _VP,
                                          _VP, _I64, C.POINTER(RenderOutputs), _VP, _I64, _VP]),
    'aninerf_render_rays_tiled': (_I32, [_VP, C.POINTER(Frame), C.POINTER(RenderParams), C.POINTER(Silhouettes), C.POINTER(PeerGather), _VP, _VP, _VP,
                                         _VP, _VP, _VP, _I64, C.POINTER(RenderOutputs), _VP, _I64, _VP]),
    'aninerf_inside_all_views': (_I32, [_VP, _I64, C.POINTER(Silhouettes), _VP, _VP]),
    'aninerf_front_end_workspace_bytes': (_I64, [_I64, _I32, _I64]),
    'aninerf_front_end': (_I32, [C.POINTER(Frame), C.POINTER(RenderParams), _VP, _VP, _VP, _VP, _VP, _VP, _I64, _VP, _VP, _VP, _VP, _VP, _VP, _VP,
                                 _VP, _I64, _VP]),
    'aninerf_gemm_workspace_bytes': (_I64, [C.POINTER(Gemm)]),
    'aninerf_gemm_x3': (_I32, [C.POINTER(Gemm), _VP, _I64, _VP]),
    'aninerf_colsum': (_I32, [_VP, _I64, _I64, _I32, _VP, _I32, _VP, _I64, _VP]),
    'aninerf_pe_forward': (_I32, [_VP, _I64, _I32, _VP, _I64, _VP]),
    'aninerf_pe_backward': (_I32, [_VP, _VP, _I64, _I64, _I32, _VP, _I32, _VP]),
    'aninerf_bw_softmax_forward': (_I32, [_VP, _I64, _VP, _I64, _VP, _VP]),
    'aninerf_bw_softmax_backward': (_I32, [_VP, _I64, _VP, _VP, _I64, _VP, _VP, _VP]),
    'aninerf_inverse_lbs_backward': (_I32, [_VP, _VP, _VP, _VP, _I64, _VP, _I32, _VP]),
    'aninerf_sample_blend_weights_backward': (_I32, [_VP, _I64, _VP, C.POINTER(_I32), _VP, _VP, _VP, _I32, _VP]),
    'aninerf_nerf_tail_forward': (_I32, [_VP, _VP, _VP, _VP, _VP, _VP, _I64, _VP, _VP, _VP]),
    'aninerf_nerf_tail_backward': (_I32, [_VP, _VP, _VP, _VP, _VP, _VP, _VP, _I64, _VP, _VP, _VP]),
    'aninerf_mask_sigma': (_I32, [_VP, _VP, _VP, _VP, _I64, _F, _I64, _VP, _VP]),
    'aninerf_composite_backward': (_I32, [_VP, _VP, _I64, _I32, _I32, _VP, _VP]),
    'aninerf_img_loss': (_I32, [_VP, _VP, _VP, _I64, _VP, _VP, _VP]),
    'aninerf_select_rows': (_I32, [_VP, _VP, _I32, _F, _VP, _VP, _VP]),
    'aninerf_gather_selected_rows': (_I32, [_VP, _VP, _I32, _VP, _VP, _VP, _VP, _VP, _VP]),
    'aninerf_bw_loss': (_I32, [_VP, _VP, _VP, _VP, _I64, _VP, _VP, _VP, _VP]),
    'aninerf_knn_blend_weights': (_I32, [_VP, _I64, _VP, _I32, _VP, _I32, _F, _VP, _VP, _VP]),
    'aninerf_marching_cubes_workspace_bytes': (_I64, [_I32, _I32, _I32]),
    'aninerf_marching_cubes': (_I32, [_VP, _I32, _I32, _I32, C.c_double, _VP, _I64, _VP, _I64, _VP, _VP, _I64, _VP]),
    'aninerf_query_workspace_bytes': (_I64, [_I64, _I64]),
    'aninerf_query_alpha': (_I32, [_VP, C.POINTER(Frame), _VP, _I64, _I64, _F, _I32, _I32, _VP, _VP, _VP, _I64, _VP]),
}

_lib = None


class AninerfError(RuntimeError):
    pass


def lib():
    """The loaded shared library; raises (never falls back) if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f'{LIB_PATH} is missing: build it with `python -c "import __graft_entry__ as g; g.build()"` '
                              f'or `make -C animatable_nerf_b200/csrc` -- there is no CPU fallback')
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc: int):
    if rc != 0:
        raise AninerfError(f'libaninerf_b200 error {rc}: {lib().aninerf_last_error().decode()}')


def ptr(t):
    """device (or host) data pointer of a tensor, None -> NULL"""
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_ptr(device=None):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def f32c(t: torch.Tensor) -> torch.Tensor:
    """contiguous float32 view/copy on the same device"""
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def require_cuda(t: torch.Tensor, name: str):
    if not t.is_cuda:
        raise AninerfError(f'{name} must be a CUDA tensor: the B200 path has no CPU fallback')
