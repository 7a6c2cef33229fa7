"""Checkpoint -> kernel operands.

Takes a `state_dict` with the reference's key layout (`lib/networks/bw_deform/tpose_nerf_network.py`
:12-38, 219-239, 279-294; Conv1d weights are (out, in, 1)) and produces, per field, the nine dense
layers the tcgen05 kernel consumes.  Two exact algebraic rewrites happen here, in float64 on the host
(they change rounding only):

 * the per-frame latent code is constant over a frame, so its columns fold into the bias:
   bias[idx] = b + W[:, latent cols] @ latent[idx]   (one bias row per latent index);
 * `feature_fc -> latent_fc -> view_fc` has no activation between its linear maps, so it collapses
   into one (256+27) -> 128 layer:  W = [Wv1 Wl1 Wf | Wv2],
   bias[idx] = Wv1 (Wl1 bf + Wl2 nf_latent[idx] + bl) + bv.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib

N_PE_XYZ = 63
N_PE_VIEW = 27
N_LATENT = 128


def _w(sd, key):
    w = sd[key].detach().to('cpu', torch.float64).numpy()
    return w[..., 0] if w.ndim == 3 else w


def _trunk_layers(sd, prefix, latent=None):
    """8 trunk layers -> list of (W (n_out,k_in) f64, bias_table (n_tables,n_out) f64).
    latent: (n_idx,128) f64 table when the trunk input is [PE(63), latent(128)], else None."""
    out = []
    for i in range(8):
        W = _w(sd, f'{prefix}.{i}.weight')
        b = _w(sd, f'{prefix}.{i}.bias')
        if latent is not None and i in (0, 5):
            Wl = W[:, N_PE_XYZ:N_PE_XYZ + N_LATENT]
            W = np.concatenate([W[:, :N_PE_XYZ], W[:, N_PE_XYZ + N_LATENT:]], axis=1)
            bias = b[None] + latent @ Wl.T
        else:
            bias = b[None]
        out.append((W, bias))
    return out


def fold_bw_field(sd, prefix=''):
    """Blend-weight field (`prefix=''`: Network.bw_linears/bw_fc/bw_latent, tpose_nerf_network.py:16-29;
    `prefix='novel_pose_bw.'`: BackwardBlendWeight, :279-294)."""
    latent = _w(sd, prefix + 'bw_latent.weight')
    layers = _trunk_layers(sd, prefix + 'bw_linears', latent)
    layers.append((_w(sd, prefix + 'bw_fc.weight'), _w(sd, prefix + 'bw_fc.bias')[None]))
    return layers


def fold_nerf_field(sd, prefix='tpose_human.'):
    """Canonical NeRF field (TPoseHuman, tpose_nerf_network.py:219-275)."""
    layers = _trunk_layers(sd, prefix + 'pts_linears', None)
    Wf, bf = _w(sd, prefix + 'feature_fc.weight'), _w(sd, prefix + 'feature_fc.bias')
    Wl, bl = _w(sd, prefix + 'latent_fc.weight'), _w(sd, prefix + 'latent_fc.bias')
    Wv, bv = _w(sd, prefix + 'view_fc.weight'), _w(sd, prefix + 'view_fc.bias')
    lat = _w(sd, prefix + 'nf_latent.weight')
    Wl1, Wl2 = Wl[:, :256], Wl[:, 256:]
    Wv1, Wv2 = Wv[:, :256], Wv[:, 256:]
    W8 = np.concatenate([Wv1 @ Wl1 @ Wf, Wv2], axis=1)                       # (128, 283)
    g = (Wl1 @ bf + bl)[None] + lat @ Wl2.T                                   # (n_idx, 256)
    bias8 = g @ Wv1.T + bv[None]                                              # (n_idx, 128)
    layers.append((W8, bias8))
    heads = (_w(sd, prefix + 'alpha_fc.weight'), _w(sd, prefix + 'alpha_fc.bias'),
             _w(sd, prefix + 'rgb_fc.weight'), _w(sd, prefix + 'rgb_fc.bias'))
    return layers, heads


def _f32(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def _fp(a):
    return a.ctypes.data_as(_lib.c_float_p)


class PackedNet:
    """Owns an `aninerf_net` (device-side packed tensor-core operand images) for one Network."""

    def __init__(self):
        self._h = C.c_void_p()
        _lib.check(_lib.lib().aninerf_net_create(C.byref(self._h)))
        self.loaded = set()

    @property
    def handle(self):
        return self._h

    def load_state_dict(self, sd, device=None):
        has_novel = any(k.startswith('novel_pose_bw.') for k in sd)
        with torch.cuda.device(device if device is not None else torch.cuda.current_device()):
            self._load(_lib.FIELD_BW, fold_bw_field(sd, ''), None)
            if has_novel:
                self._load(_lib.FIELD_NOVEL_BW, fold_bw_field(sd, 'novel_pose_bw.'), None)
            layers, heads = fold_nerf_field(sd)
            self._load(_lib.FIELD_NERF, layers, heads)

    def _load(self, field, layers, heads):
        n = len(layers)
        arr = (_lib.Layer * n)()
        keep = []
        for i, (W, bias) in enumerate(layers):
            Wc, bc = _f32(W), _f32(bias)
            keep += [Wc, bc]
            arr[i].W = _fp(Wc)
            arr[i].bias_table = _fp(bc)
            arr[i].n_out, arr[i].k_in = Wc.shape
            arr[i].n_tables = bc.shape[0]
            arr[i].relu = 0 if i == n - 1 and heads is None else 1
        hp = [None] * 4
        if heads is not None:
            hs = [_f32(h) for h in heads]
            keep += hs
            hp = [_fp(h) for h in hs]
        _lib.check(_lib.lib().aninerf_net_load_field(self._h, field, arr, n, hp[0], hp[1], hp[2], hp[3], _lib.stream_ptr()))
        self.loaded.add(field)

    def close(self):
        if self._h:
            _lib.lib().aninerf_net_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
