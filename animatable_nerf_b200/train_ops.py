"""ctypes wrappers of the training-step entries of libaninerf_b200.so (include/aninerf_b200.h, "Training step").

Operands are described in place: `Op(t, rows, k)` views a 2-D fp32 CUDA tensor (or a column range of one) as
`rows x k` with element (r, k) at `t[r, k]`; `.T` swaps the roles, so X, X^T, W, W^T and W[:, a:b] are all read
without copies.  No CPU fallback.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib


class Op:
    """element (r, k) = base[r * row_stride + k * k_stride], r < rows, k < K"""
    __slots__ = ('t', 'ptr', 'rows', 'k', 'rs', 'ks')

    def __init__(self, t: torch.Tensor, transpose: bool = False):
        assert t.is_cuda and t.dtype == torch.float32 and t.dim() == 2, 'Op needs a 2-D fp32 CUDA tensor'
        self.t = t
        self.ptr = t.data_ptr()
        if transpose:
            self.rows, self.k, self.rs, self.ks = t.shape[1], t.shape[0], t.stride(1), t.stride(0)
        else:
            self.rows, self.k, self.rs, self.ks = t.shape[0], t.shape[1], t.stride(0), t.stride(1)

    @property
    def T(self):
        o = Op.__new__(Op)
        o.t, o.ptr, o.rows, o.k, o.rs, o.ks = self.t, self.ptr, self.k, self.rows, self.ks, self.rs
        return o


class Workspace:
    """Grow-only scratch for split-K partials / column sums."""

    def __init__(self):
        self.buf = None

    def get(self, nbytes, device):
        if self.buf is None or self.buf.numel() < nbytes or self.buf.device != device:
            self.buf = torch.empty(max(int(nbytes), 1 << 20), dtype=torch.uint8, device=device)
        return self.buf


_ws = Workspace()


_GEMM = None        # the library entry, resolved once (ctypes attribute lookups are slow on a ~240-launch path)
_stream = {}        # device -> c_void_p of the step's stream (set by begin_step; torch.cuda.current_stream costs microseconds per call)


def begin_step(device):
    """Cache the current stream of `device` for the wrappers below; call once at the start of a training step."""
    _stream[device] = _lib.stream_ptr(device)


def end_step(device):
    _stream.pop(device, None)


def _st(device):
    s = _stream.get(device)
    return s if s is not None else _lib.stream_ptr(device)


def gemm(segs, out: torch.Tensor, bias=None, relu=False, relu_mask=None, accumulate=False, split_k=1):
    """out[M,N] = epilogue(sum_s A_s @ B_s^T); segs = [(Op A (M x K), Op B (N x K)), ...]; out may be a strided 2-D view."""
    global _GEMM
    if _GEMM is None:
        _GEMM = _lib.lib().aninerf_gemm_x3
    g = _lib.Gemm()
    g.n_seg = len(segs)
    M, N = out.shape
    for i, (a, b) in enumerate(segs):
        assert a.rows == M and b.rows == N and a.k == b.k, (a.rows, a.k, b.rows, b.k, M, N)
        s = g.seg[i]
        s.A, s.a_row_stride, s.a_k_stride = a.ptr, a.rs, a.ks
        s.B, s.b_row_stride, s.b_k_stride = b.ptr, b.rs, b.ks
        s.K = a.k
    g.M, g.N, g.C, g.ldc = M, N, out.data_ptr(), out.stride(0)
    if bias is not None:
        g.bias = bias.data_ptr()
    if relu_mask is not None:
        g.relu_mask, g.ld_mask = relu_mask.data_ptr(), relu_mask.stride(0)
    g.relu, g.accumulate, g.split_k = int(relu), int(accumulate), int(split_k)
    ws, nbytes = None, 0
    if split_k > 1:
        nbytes = split_k * M * N * 4            # = aninerf_gemm_workspace_bytes
        ws = _ws.get(nbytes, out.device)
    rc = _GEMM(C.byref(g), _lib.ptr(ws), nbytes, _st(out.device))
    if rc:
        _lib.check(rc)
    return out


def split_for(k: int) -> int:
    """split-K factor of a weight-gradient product reducing over k samples"""
    return max(1, min(64, k // 128))


def colsum(x: torch.Tensor, out: torch.Tensor, accumulate=False):
    assert x.dim() == 2 and x.stride(1) == 1 and out.numel() == x.shape[1] and out.is_contiguous()
    M, N = x.shape
    nbytes = ((M + 255) // 256) * N * 4
    ws = _ws.get(nbytes, x.device)
    _lib.check(_lib.lib().aninerf_colsum(_lib.ptr(x), x.stride(0), M, N, _lib.ptr(out), int(accumulate), _lib.ptr(ws), nbytes,
                                         _st(x.device)))
    return out


def pe_forward(x, n_freq, out):
    _lib.check(_lib.lib().aninerf_pe_forward(_lib.ptr(x), x.shape[0], n_freq, _lib.ptr(out), out.stride(0), _st(x.device)))
    return out


def pe_backward(x, d_pe, n_freq, d_x, accumulate):
    _lib.check(_lib.lib().aninerf_pe_backward(_lib.ptr(x), _lib.ptr(d_pe), d_pe.stride(0), x.shape[0], n_freq, _lib.ptr(d_x), int(accumulate),
                                              _st(x.device)))
    return d_x


def bw_softmax_forward(init, delta, bw):
    _lib.check(_lib.lib().aninerf_bw_softmax_forward(_lib.ptr(init), init.stride(0), _lib.ptr(delta), delta.shape[0], _lib.ptr(bw),
                                                     _st(bw.device)))
    return bw


def bw_softmax_backward(init, bw, d_bw, d_delta, d_init=None):
    _lib.check(_lib.lib().aninerf_bw_softmax_backward(_lib.ptr(init), init.stride(0), _lib.ptr(bw), _lib.ptr(d_bw), bw.shape[0], _lib.ptr(d_delta),
                                                      _lib.ptr(d_init), _st(bw.device)))


def inverse_lbs(ppts, bw, A, tpts):
    _lib.check(_lib.lib().aninerf_inverse_lbs(_lib.ptr(ppts), _lib.ptr(bw), ppts.shape[0], _lib.ptr(A), _lib.ptr(tpts), _st(ppts.device)))
    return tpts


def inverse_lbs_backward(bw, A, tpts, d_tpts, d_bw, accumulate):
    _lib.check(_lib.lib().aninerf_inverse_lbs_backward(_lib.ptr(bw), _lib.ptr(A), _lib.ptr(tpts), _lib.ptr(d_tpts), bw.shape[0], _lib.ptr(d_bw),
                                                       int(accumulate), _st(bw.device)))


def sample_volume(pts, vol, bounds, out25):
    dims = (C.c_int32 * 3)(*vol.shape[-4:-1])
    _lib.check(_lib.lib().aninerf_sample_blend_weights(_lib.ptr(pts), pts.shape[0], _lib.ptr(vol), dims, _lib.ptr(bounds), _lib.ptr(out25),
                                                       _st(pts.device)))
    return out25


def sample_volume_backward(pts, vol, bounds, d_out24, d_pts, accumulate):
    dims = (C.c_int32 * 3)(*vol.shape[-4:-1])
    _lib.check(_lib.lib().aninerf_sample_blend_weights_backward(_lib.ptr(pts), pts.shape[0], _lib.ptr(vol), dims, _lib.ptr(bounds), _lib.ptr(d_out24),
                                                                _lib.ptr(d_pts), int(accumulate), _st(pts.device)))


def world_to_pose(wpts, R, Th, out):
    _lib.check(_lib.lib().aninerf_world_to_pose(_lib.ptr(wpts), wpts.shape[0], _lib.ptr(R), _lib.ptr(Th), _lib.ptr(out), _st(wpts.device)))
    return out


def forward_lbs(tpts, bw, A, ppts):
    _lib.check(_lib.lib().aninerf_forward_lbs(_lib.ptr(tpts), _lib.ptr(bw), tpts.shape[0], _lib.ptr(A), _lib.ptr(ppts), _st(tpts.device)))
    return ppts


def mask_sigma(sigma, tpts, tbounds, pnorm, ld, norm_th, out):
    _lib.check(_lib.lib().aninerf_mask_sigma(_lib.ptr(sigma), _lib.ptr(tpts), _lib.ptr(tbounds), _lib.ptr(pnorm), ld, float(norm_th), tpts.shape[0],
                                             _lib.ptr(out), _st(tpts.device)))
    return out


def select_rows(sigma_masked, chunk_offsets, n_chunks, train_th, sel, n_sel):
    _lib.check(_lib.lib().aninerf_select_rows(_lib.ptr(sigma_masked), _lib.ptr(chunk_offsets), n_chunks, float(train_th), _lib.ptr(sel), _lib.ptr(n_sel),
                                              _st(sel.device)))


def bw_loss(pbw, tbw, sel, n_sel, loss, d_pbw, d_tbw):
    _lib.check(_lib.lib().aninerf_bw_loss(_lib.ptr(pbw), _lib.ptr(tbw), _lib.ptr(sel), _lib.ptr(n_sel), pbw.shape[0], _lib.ptr(loss), _lib.ptr(d_pbw),
                                          _lib.ptr(d_tbw), _st(pbw.device)))
