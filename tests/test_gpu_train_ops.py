"""GPU numerics of the training-step kernels through the C ABI: the split-precision tcgen05 GEMM and the
per-point forward/backward ops, each against a plain PyTorch reference of the same op (fp64 for the GEMM)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from helpers import O, small_frame_case

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def dev():
    assert torch.cuda.is_available(), 'these tests need the B200'
    return torch.device('cuda:0')


def _rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


GEMM_TOL = 4e-5      # bf16x3: hi+lo keeps ~17 mantissa bits per operand -> ~2^-16 relative per product; fp32 accumulation


@pytest.mark.parametrize('M,N,K', [(1, 256, 128), (300, 256, 63), (4099, 256, 256), (777, 24, 256), (513, 1, 256), (129, 3, 128), (1000, 128, 283),
                                   (256, 191, 5000)])
def test_gemm_plain_and_transposed_operands(dev, M, N, K):
    from animatable_nerf_b200 import train_ops as T
    g = torch.Generator().manual_seed(M + N + K)
    a = torch.randn(M, K, generator=g).to(dev)
    b = torch.randn(N, K, generator=g).to(dev)
    ref = a.double() @ b.double().t()
    out = torch.full((M, N), float('nan'), device=dev)
    T.gemm([(T.Op(a), T.Op(b))], out)
    assert _rel(out, ref) <= GEMM_TOL
    # both operands stored transposed (row-contiguous staging path)
    at, bt = a.t().contiguous(), b.t().contiguous()
    out2 = torch.full((M, N), float('nan'), device=dev)
    T.gemm([(T.Op(at).T, T.Op(bt).T)], out2)
    assert _rel(out2, ref) <= GEMM_TOL
    # split-K (weight-gradient form) with accumulation onto an existing value
    base = torch.randn(M, N, generator=g).to(dev)
    out3 = base.clone()
    T.gemm([(T.Op(a), T.Op(b))], out3, accumulate=True, split_k=7)
    assert _rel(out3, ref + base.double()) <= GEMM_TOL


def test_gemm_segments_and_epilogues(dev):
    from animatable_nerf_b200 import train_ops as T
    g = torch.Generator().manual_seed(1)
    M = 1500
    pe = torch.randn(M, 63, generator=g).to(dev)
    hid = torch.randn(M, 256, generator=g).to(dev)
    W = (torch.randn(256, 447, generator=g) / 16).to(dev)          # skip layer of the blend-weight field: [PE | latent | hidden]
    bias = torch.randn(256, generator=g).to(dev)
    ref = F.relu(pe.double() @ W[:, :63].double().t() + hid.double() @ W[:, 191:].double().t() + bias.double())
    wide = torch.zeros(M, 320, device=dev)                         # the output lands in a column range of a wider buffer
    out = wide[:, 64:]
    T.gemm([(T.Op(pe), T.Op(W[:, :63])), (T.Op(hid), T.Op(W[:, 191:]))], out, bias=bias, relu=True)
    assert _rel(out, ref) <= GEMM_TOL
    assert float(wide[:, :64].abs().max()) == 0.0
    # data gradient with the ReLU mask of the producing layer and accumulation of a second path
    dz = torch.randn(M, 256, generator=g).to(dev)
    first = torch.randn(M, 256, generator=g).to(dev)
    ref_d = (first.double() + dz.double() @ W[:, 191:].double()) * (hid.double() > 0)
    dst = first.clone()
    T.gemm([(T.Op(dz), T.Op(W[:, 191:]).T)], dst, accumulate=True, relu_mask=hid)
    assert _rel(dst, ref_d) <= GEMM_TOL
    # weight gradient into a column range of the full gradient matrix + bias gradient
    dW = torch.zeros(256, 447, device=dev)
    T.gemm([(T.Op(dz).T, T.Op(hid).T)], dW[:, 191:], split_k=T.split_for(M))
    assert _rel(dW[:, 191:], dz.double().t() @ hid.double()) <= GEMM_TOL
    assert float(dW[:, :191].abs().max()) == 0.0
    db = torch.ones(256, device=dev)
    T.colsum(dz, db, accumulate=True)
    assert _rel(db, 1 + dz.double().sum(0)) <= 1e-6


def test_pointwise_ops_forward_backward(dev):
    from animatable_nerf_b200 import train_ops as T
    frame, _, batch, _ = small_frame_case(voxel=0.05, H=96, W=96, focal=100.0)
    g = torch.Generator().manual_seed(5)
    n = 5000
    lo, hi = batch['tbounds'][0, 0], batch['tbounds'][0, 1]
    pts = (torch.rand(n, 3, generator=g) * (hi - lo) * 1.1 + lo - 0.05 * (hi - lo))      # some points outside (border clipping)
    # positional encoding
    x = pts.clone().requires_grad_(True)
    pe_ref = O.positional_encoding(x, 10)
    gpe = torch.randn(n, 63, generator=g)
    pe_ref.backward(gpe)
    out = torch.empty(n, 64, device=dev)
    T.pe_forward(pts.to(dev), 10, out)
    assert float((out[:, :63].cpu() - pe_ref.detach()).abs().max()) <= 2e-6
    dx = torch.empty(n, 3, device=dev)
    gp = torch.zeros(n, 64)
    gp[:, :63] = gpe
    T.pe_backward(pts.to(dev), gp.to(dev), 10, dx, False)
    assert _rel(dx.cpu(), x.grad) <= 1e-5
    # trilinear sampling: coordinate gradient
    x = pts.clone().requires_grad_(True)
    s_ref = O.sample_blend_weights(x[None], batch['tbw'], batch['tbounds'])[0, :24].t()
    gs = torch.randn(n, 24, generator=g)
    s_ref.backward(gs)
    dp = torch.empty(n, 3, device=dev)
    T.sample_volume_backward(pts.to(dev), batch['tbw'][0].to(dev), batch['tbounds'][0].to(dev), gs.to(dev), dp, False)
    assert _rel(dp.cpu(), x.grad) <= 1e-5
    # softmax head
    init = torch.softmax(torch.randn(n, 25, generator=g) * 3, dim=1)
    delta = torch.randn(n, 24, generator=g)
    i_r, d_r = init.clone().requires_grad_(True), delta.clone().requires_grad_(True)
    bw_ref = torch.softmax(torch.log(i_r[:, :24] + 1e-9) + d_r, dim=1)
    gb = torch.randn(n, 24, generator=g)
    bw_ref.backward(gb)
    bw = torch.empty(n, 24, device=dev)
    T.bw_softmax_forward(init.to(dev), delta.to(dev), bw)
    assert float((bw.cpu() - bw_ref.detach()).abs().max()) <= 1e-6
    dd, di = torch.empty(n, 24, device=dev), torch.empty(n, 24, device=dev)
    T.bw_softmax_backward(init.to(dev), bw, gb.to(dev), dd, di)
    assert _rel(dd.cpu(), d_r.grad) <= 1e-5
    assert _rel(di.cpu(), i_r.grad[:, :24]) <= 1e-4
    # inverse LBS
    ppts = pts
    w = torch.softmax(torch.randn(n, 24, generator=g), dim=1)
    w_r = w.clone().requires_grad_(True)
    t_ref = O.inverse_lbs(ppts[None], w_r.t()[None], batch['A'])[0]
    gt = torch.randn(n, 3, generator=g)
    t_ref.backward(gt)
    A = batch['A'][0].to(dev)
    tp = torch.empty(n, 3, device=dev)
    T.inverse_lbs(ppts.to(dev), w.to(dev), A, tp)
    assert float((tp.cpu() - t_ref.detach()).abs().max()) <= 1e-5
    dw = torch.empty(n, 24, device=dev)
    T.inverse_lbs_backward(w.to(dev), A, tp, gt.to(dev), dw, False)
    assert _rel(dw.cpu(), w_r.grad) <= 1e-4


def test_composite_backward(dev):
    from animatable_nerf_b200 import _lib
    g = torch.Generator().manual_seed(9)
    R, S = 999, 64
    raw = torch.rand(R, S, 4, generator=g)
    raw[..., 3] *= (torch.rand(R, S, generator=g) < 0.3)          # mostly empty space, as in a frame
    z = torch.sort(torch.rand(R, S, generator=g) + 2, dim=1)[0]
    for white in (False, True):
        r = raw.double().requires_grad_(True)                      # fp64 autograd of the oracle's raw2outputs
        rgb_map = O.raw2outputs(r, z.double(), white)[0]
        gmap = torch.randn(R, 3, generator=g)
        rgb_map.backward(gmap.double())
        d_raw = torch.empty(R, S, 4, device=dev)
        raw_d, gmap_d = raw.to(dev), gmap.to(dev)                  # keep the device copies alive across the launch
        _lib.check(_lib.lib().aninerf_composite_backward(_lib.ptr(raw_d), _lib.ptr(gmap_d), R, S, int(white), _lib.ptr(d_raw),
                                                         _lib.stream_ptr(dev)))
        assert _rel(d_raw.cpu(), r.grad) <= 1e-5


def test_select_and_gather_rows_match_per_chunk_loop(dev):
    """aninerf_select_rows + aninerf_gather_selected_rows == `alpha_ind` / `pbw[alpha_ind]` of tpose_nerf_network.py:192-196
    evaluated chunk by chunk: threshold, first arg-max forced (ties, single-row and all-below-threshold chunks), ascending rows."""
    import ctypes as C
    from animatable_nerf_b200 import _lib
    L = _lib.lib()
    g = torch.Generator().manual_seed(2)
    counts = torch.tensor([5, 1, 40, 7, 3, 0, 700, 2048, 1])
    off = torch.cat([torch.zeros(1, dtype=torch.long), counts.cumsum(0)]).int()
    n = int(counts.sum())
    sig = torch.randn(n, generator=g)
    sig[5] = -3.0                      # single-row chunk below the threshold: must still be kept
    sig[6:46] = -1.0                   # a whole chunk below threshold with ties: first row wins
    want = sig > 0
    for c in range(len(counts)):
        a, b = int(off[c]), int(off[c + 1])
        if b > a:
            want[a + int(torch.argmax(sig[a:b]))] = True
    a24, b24 = torch.randn(n, 24, generator=g), torch.randn(n, 24, generator=g)
    d_sig, d_off, d_a, d_b = sig.to(dev), off.to(dev), a24.to(dev), b24.to(dev)
    sel = torch.zeros(n, dtype=torch.uint8, device=dev)
    n_sel = torch.zeros(1, dtype=torch.int32, device=dev)
    st = _lib.stream_ptr(dev)
    _lib.check(L.aninerf_select_rows(_lib.ptr(d_sig), _lib.ptr(d_off), len(counts), 0.0, _lib.ptr(sel), _lib.ptr(n_sel), st))
    assert torch.equal(sel.bool().cpu(), want) and int(n_sel.item()) == int(want.sum())
    k = int(want.sum())
    o_a, o_b = torch.empty(k, 24, device=dev), torch.empty(k, 24, device=dev)
    offs = torch.empty(len(counts) + 1, dtype=torch.int32, device=dev)
    _lib.check(L.aninerf_gather_selected_rows(_lib.ptr(sel), _lib.ptr(d_off), len(counts), _lib.ptr(d_a), _lib.ptr(d_b), _lib.ptr(o_a), _lib.ptr(o_b),
                                              _lib.ptr(offs), st))
    assert torch.equal(o_a.cpu(), a24[want]) and torch.equal(o_b.cpu(), b24[want])
    assert int(offs[-1].item()) == k
    per_chunk = [int(want[int(off[c]):int(off[c + 1])].sum()) for c in range(len(counts))]
    assert offs[:-1].cpu().tolist() == [sum(per_chunk[:c]) for c in range(len(counts))]
