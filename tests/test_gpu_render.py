"""GPU parity of the fused path: Renderer.render(batch) through aninerf_render_rays vs the CPU oracle
and vs the committed reference outputs (tests/golden)."""
import numpy as np
import pytest
import torch

from helpers import O, check_selected_rows, golden_small_case, load_golden, small_frame_case, to_device

pytestmark = pytest.mark.gpu

BW_TOL = 1e-5
RGB_TOL = 2e-3


@pytest.fixture(scope='module')
def dev():
    assert torch.cuda.is_available(), 'these tests need the B200'
    return torch.device('cuda:0')

@pytest.fixture(autouse=True)
def _no_grad():
    """Evaluation runs under torch.no_grad() (run.py:62 of the reference); with gradients enabled Renderer.render returns device
    tensors carrying the graph (tests/test_gpu_train.py covers that mode)."""
    with torch.no_grad():
        yield


def _renderer(dev, sd, **over):
    from animatable_nerf_b200 import config
    from animatable_nerf_b200.tpose_nerf_network import Network
    from animatable_nerf_b200.tpose_renderer import Renderer
    cfg = config.make_cfg(perturb=0., **over)
    net = Network(cfg)
    net.load_state_dict(sd)
    net = net.to(dev)
    return Renderer(net, cfg)


def _check_maps(out, ref, tol=RGB_TOL):
    for k in ('rgb_map', 'acc_map', 'depth_map'):
        assert out[k].shape == ref[k].shape, k
        assert float((out[k].cpu() - torch.as_tensor(ref[k])).abs().max()) <= tol, k


def test_render_matches_reference_golden(dev):
    """The committed REFERENCE outputs (2175 rays = one full 2048-ray chunk + a ragged one)."""
    g, batch, sd = golden_small_case()
    r = _renderer(dev, sd)
    out = r.render(to_device(batch, dev))
    assert set(out) == {'rgb_map', 'acc_map', 'depth_map', 'raw', 'pbw', 'tbw'}
    assert all(not v.is_cuda for v in out.values())            # tpose_renderer.py:154-155
    _check_maps(out, g)
    assert out['raw'].shape == g['raw'].shape
    assert float((out['raw'] - torch.from_numpy(g['raw'])).abs().max()) <= RGB_TOL
    # the zero pattern of raw is the (bit-exact) pnorm mask
    assert np.array_equal((out['raw'].numpy()[..., :3] != 0).any(-1), (g['raw'][..., :3] != 0).any(-1))


def test_render_internals_vs_oracle(dev):
    """Active set bit-exact; canonical points and both blend-weight fields within 1e-5; the pbw/tbw
    row selection (alpha > train_th + per-chunk arg-max) reproduces the oracle's rows."""
    _, _, batch, _ = small_frame_case(voxel=0.05, H=160, W=160, focal=170.0)
    from animatable_nerf_b200 import synthetic
    sd = synthetic.make_state_dict(seed=0)
    ref = O.render(sd, batch, O.OracleCfg(perturb=0.), return_debug=True)
    dbg = ref['_debug']
    r = _renderer(dev, sd)
    dv = r.render_device(to_device(batch, dev), want_bw=True)
    n_active = int(dv['n_active'].item())
    pind = dbg['pind'].numpy()
    assert n_active == int(pind.sum())
    assert np.array_equal(dv['active_index'][:n_active].cpu().numpy(), np.nonzero(pind)[0].astype(np.int32))
    assert np.array_equal(np.diff(dv['chunk_offsets'].cpu().numpy()), dbg['chunk_active'].numpy())
    assert float((dv['pbw_all'][:n_active].cpu() - dbg['pbw_all']).abs().max()) <= BW_TOL
    assert float((dv['tbw_all'][:n_active].cpu() - dbg['tbw_all']).abs().max()) <= BW_TOL
    _check_maps({k: dv[k].view(ref[k].shape) for k in ('rgb_map', 'acc_map', 'depth_map')}, ref)
    # `outside` (canonical tbounds test) sits behind the MLP: mismatches only for points within 1e-5 of a bound
    sig = dv['sigma_masked'][:n_active].cpu()
    out_gpu = sig == 0
    mism = (out_gpu != dbg['outside']).nonzero().reshape(-1)
    tb = batch['tbounds'][0]
    for i in mism.tolist():
        margin = torch.minimum((dbg['tpose'][i] - tb[0]).abs().min(), (dbg['tpose'][i] - tb[1]).abs().min())
        assert float(margin) <= 1e-5, f'outside mismatch at row {i} with margin {float(margin)}'
    # the pbw / tbw rows of the contract: row SET vs the oracle (mismatches only within the sigma noise of the threshold /
    # of the chunk maximum), values on the common rows within 1e-5 -- unconditionally
    rows, cg, ref_rows, cr, pbw, tbw, n_mism, n_near = check_selected_rows(r, dv, dbg)
    assert n_mism <= n_near
    assert float((pbw[cg] - ref['pbw'][0][cr]).abs().max()) <= BW_TOL
    assert float((tbw[cg] - ref['tbw'][0][cr]).abs().max()) <= BW_TOL
    # full contract through render(): the same rows, on the host
    out = r.render(to_device(batch, dev))
    assert out['pbw'].shape == (1, rows.numel(), 24) and out['tbw'].shape == out['pbw'].shape
    assert torch.equal(out['pbw'][0], pbw) and torch.equal(out['tbw'][0], tbw)
    assert not out['pbw'].is_cuda and not out['raw'].is_cuda


def test_render_with_jitter(dev):
    g, batch, sd = golden_small_case()
    r = _renderer(dev, sd)
    t_rand = torch.from_numpy(g['jitter_t_rand'])
    dv = r.render_device(to_device(batch, dev), t_rand=t_rand.to(dev), want_bw=False)
    _check_maps({k: dv[k].view(g['jitter_' + k].shape) for k in ('rgb_map', 'acc_map', 'depth_map')},
                {k: g['jitter_' + k] for k in ('rgb_map', 'acc_map', 'depth_map')})


def test_render_novel_pose_field(dev):
    from animatable_nerf_b200 import synthetic
    g, batch, _ = golden_small_case()
    g2 = load_golden('render_small_novel_pose.npz')
    sd2 = synthetic.make_state_dict(seed=int(g2['sd_seed']), num_eval_frame=int(g2['num_eval_frame']))
    r = _renderer(dev, sd2, aninerf_animation=True, test_novel_pose=True, num_eval_frame=int(g2['num_eval_frame']))
    out = r.render(to_device(batch, dev))
    _check_maps(out, g2)


def test_render_only_mode_and_edge_sizes(dev):
    g, batch, sd = golden_small_case()
    r = _renderer(dev, sd, b200_render_only=True)
    out = r.render(to_device(batch, dev))
    assert set(out) == {'rgb_map', 'acc_map', 'depth_map'}
    _check_maps(out, g)
    for n in (1, 63, 2048, 2049):
        sub = {k: (v[:, :n] if k in ('ray_o', 'ray_d', 'near', 'far', 'occupancy') else v) for k, v in batch.items()}
        ref = O.render(sd, sub, O.OracleCfg(perturb=0.))
        _check_maps(r.render(to_device(sub, dev)), ref)
    # render-only composites the compact active rows (no dense (n,4) buffer): bit-identical to compositing the dense raw
    b = to_device(batch, dev)
    compact = r.render_device(b, want_bw=False)
    dense = r.render_device(b, want_bw=False, keep_raw=True)
    assert 'raw' not in compact and 'raw' in dense
    for k in ('rgb_map', 'acc_map', 'depth_map'):
        assert torch.equal(compact[k], dense[k]), k


def test_chunk_with_no_active_sample_forces_argmin(dev):
    """Rays that miss the SMPL shell entirely: the reference still evaluates one sample per chunk."""
    _, batch, sd = golden_small_case()
    sub = {k: (v[:, :300].clone() if k in ('ray_o', 'ray_d', 'near', 'far', 'occupancy') else v) for k, v in batch.items()}
    sub['near'] = sub['near'] * 0 + 50.0
    sub['far'] = sub['far'] * 0 + 51.0
    ref = O.render(sd, sub, O.OracleCfg(perturb=0.), return_debug=True)
    assert int(ref['_debug']['pind'].sum()) == 1
    r = _renderer(dev, sd)
    dv = r.render_device(to_device(sub, dev), want_bw=True)
    assert int(dv['n_active'].item()) == 1
    assert int(dv['active_index'][0].item()) == int(ref['_debug']['pind'].nonzero()[0, 0])
    _check_maps({k: dv[k].view(ref[k].shape) for k in ('rgb_map', 'acc_map', 'depth_map')}, ref)


def test_density_query(dev):
    g, batch, sd = golden_small_case()
    from animatable_nerf_b200 import config
    from animatable_nerf_b200.tpose_nerf_network import Network
    net = Network(config.make_cfg())
    net.load_state_dict(sd)
    net = net.to(dev)
    pts, _ = O.sample_points(batch['ray_o'], batch['ray_d'], batch['near'], batch['far'], 64)
    w = pts.view(-1, 3)[:20000]
    got = net.get_alpha(w.to(dev), to_device(batch, dev))
    ref = torch.from_numpy(g['alpha_grid'])
    assert np.array_equal((got.cpu() != 0).numpy(), (ref != 0).numpy())      # mask (norm_th 0.1) bit-exact
    assert float((got.cpu() - ref).abs().max()) <= 5e-3


def test_network_forward_api(dev):
    """Network.forward(wpts, viewdir, dists, batch) for one explicit chunk, as the reference renderer calls it."""
    _, batch, sd = golden_small_case()
    sub = {k: (v[:, :500] if k in ('ray_o', 'ray_d', 'near', 'far', 'occupancy') else v) for k, v in batch.items()}
    pts, z = O.sample_points(sub['ray_o'], sub['ray_d'], sub['near'], sub['far'], 64)
    wpts = pts.view(-1, 3)
    vd = sub['ray_d'][0][:, None].repeat(1, 64, 1).view(-1, 3)
    dists = O.sample_dists(z).view(-1)
    ref = O.network_forward(sd, wpts, vd, dists, sub, O.OracleCfg())
    r = _renderer(dev, sd)
    out = r.net(wpts.to(dev), vd.to(dev), dists.to(dev), to_device(sub, dev))
    assert out['raw'].shape == ref['raw'].shape
    assert float((out['raw'].cpu() - ref['raw']).abs().max()) <= RGB_TOL
    # one chunk: the selected rows are sigma > 0 plus the arg-max; the row count may differ only by rows within the bf16 sigma noise of 0
    refd = O.network_forward(sd, wpts, vd, dists, sub, O.OracleCfg(), return_debug=True)['_debug']
    near_zero = int((refd['sigma_masked'].abs() <= 5e-3).sum())
    assert abs(out['pbw'].shape[1] - ref['pbw'].shape[1]) <= near_zero
    if near_zero == 0:
        assert float((out['pbw'].cpu() - ref['pbw']).abs().max()) <= BW_TOL and float((out['tbw'].cpu() - ref['tbw']).abs().max()) <= BW_TOL


# ---------------------------------------------------------------------------------------------
# tpose_renderer_mmsk: multi-view silhouette culling ahead of the network
# ---------------------------------------------------------------------------------------------
def _mmsk_renderer(dev, sd, **over):
    from animatable_nerf_b200 import config
    from animatable_nerf_b200.tpose_nerf_network import Network
    from animatable_nerf_b200.tpose_renderer_mmsk import Renderer
    cfg = config.make_cfg(perturb=0., **over)
    net = Network(cfg)
    net.load_state_dict(sd)
    return Renderer(net.to(dev), cfg)


def _golden_mmsk_batch():
    g, batch, sd = golden_small_case()
    gm = load_golden('render_small_mmsk.npz')
    mb = dict(batch)
    mb['msks'] = torch.from_numpy(gm['msks'])[None]
    mb['Ks'] = torch.from_numpy(gm['Ks'])[None]
    mb['RT'] = torch.from_numpy(gm['RT'])[None]
    mb['H'], mb['W'] = torch.tensor([int(gm['H'])]), torch.tensor([int(gm['W'])])
    return gm, mb, sd


def test_inside_all_views_bit_exact(dev):
    """prepare_inside_pts through aninerf_inside_all_views vs the REFERENCE's mask (golden) and the oracle
    on a larger random cloud (incl. points behind cameras / far outside the images)."""
    gm, mb, sd = _golden_mmsk_batch()
    r = _mmsk_renderer(dev, sd)
    pts, _ = O.sample_points(mb['ray_o'], mb['ray_d'], mb['near'], mb['far'], 64)
    got = r.prepare_inside_pts(pts.to(dev), to_device(mb, dev))
    assert got.shape == (1, pts.shape[1] * 64)
    assert np.array_equal(got.cpu().numpy().reshape(-1), gm['inside'])
    g = torch.Generator().manual_seed(7)
    cloud = (torch.rand(1, 400000, 1, 3, generator=g) - 0.5) * 8.0
    ref = O.inside_all_views(cloud.view(1, -1, 3), mb)
    got = r.prepare_inside_pts(cloud.to(dev), to_device(mb, dev))
    assert torch.equal(got.cpu(), ref)
    assert 0 < int(ref.sum()) < ref.numel()


def test_mmsk_render_matches_reference_golden(dev):
    gm, mb, sd = _golden_mmsk_batch()
    r = _mmsk_renderer(dev, sd)
    out = r.render(to_device(mb, dev))
    assert set(out) == {'rgb_map', 'acc_map', 'depth_map'}          # tpose_renderer_mmsk.py:135-139
    assert all(not v.is_cuda for v in out.values())
    _check_maps(out, gm)
    # culled + masked active set is bit-exact: compare the zero pattern of raw with the oracle's
    dev_out = r.render_device(to_device(mb, dev))
    ref = O.render_mmsk(sd, mb, O.OracleCfg(perturb=0.), return_debug=True)
    assert int(dev_out['n_active'].item()) == int(ref['_debug']['chunk_active'].sum())


def test_mmsk_chunk_without_survivors_evaluates_nothing(dev):
    """All-zero silhouettes: no sample reaches the network (no argmin forcing either), maps are the empty render."""
    gm, mb, sd = _golden_mmsk_batch()
    mb = dict(mb)
    mb['msks'] = torch.zeros_like(mb['msks'])
    r = _mmsk_renderer(dev, sd)
    dev_out = r.render_device(to_device(mb, dev))
    assert int(dev_out['n_active'].item()) == 0
    assert float(dev_out['acc_map'].abs().max()) == 0.0 and float(dev_out['rgb_map'].abs().max()) == 0.0
    ref = O.render_mmsk(sd, mb, O.OracleCfg(perturb=0.))
    _check_maps({k: dev_out[k][None] for k in ('rgb_map', 'acc_map', 'depth_map')}, ref)


def test_full_frame_properties_at_baseline_size(dev):
    """BASELINE config 2 (1024x1024, ~243 k rays x 64): too big for the oracle, so size-independent properties --
    (1) ray-tile invariance: the chunks 0,2,4,.. and 1,3,5,.. rendered separately and re-interleaved equal the whole-frame
        render bit for bit (what the N-GPU path relies on);
    (2) compact compositing == dense compositing bit for bit; the dense raw is zero exactly off the active set;
    (3) ranges: acc in [0,1], rgb in [0,1], depth within [0, max far]; rays whose 64 samples are all inactive composite to 0."""
    from animatable_nerf_b200 import frontend, ray_tiles, synthetic
    frame = synthetic.make_frame(pose_seed=2, body_seed=1, voxel=0.025)
    K, R, T = synthetic.make_camera(frame, 1024, 1024)
    ro, rd, near, far, mask = frontend.get_rays_within_bounds(1024, 1024, K, R, T, frame['wbounds'], device=dev)
    n = ro.shape[0]
    assert 200_000 < n < 300_000
    sd = synthetic.make_state_dict(seed=0)
    r = _renderer(dev, sd, b200_render_only=True)
    full = synthetic.make_render_batch(frame, ro, rd, near, far, device=dev)
    whole = r.render_device(full, want_bw=False, keep_raw=True)
    maps = torch.cat([whole['rgb_map'], whole['acc_map'][:, None], whole['depth_map'][:, None]], dim=1)
    tiled = torch.empty_like(maps)
    for rank in range(2):
        idx = ray_tiles.shard_indices(n, rank, 2, device=dev)
        part = r.render_device(ray_tiles.shard_batch(full, rank, 2), want_bw=False)
        tiled[idx] = torch.cat([part['rgb_map'], part['acc_map'][:, None], part['depth_map'][:, None]], dim=1)
    assert torch.equal(tiled, maps)
    compact = r.render_device(full, want_bw=False)
    for k in ('rgb_map', 'acc_map', 'depth_map'):
        assert torch.equal(compact[k], whole[k]), k
    na = int(whole['n_active'].item())
    raw = whole['raw'].view(n, 64, 4)
    assert int((raw[..., :3] != 0).any(-1).sum()) == na            # sigmoid(rgb) > 0 exactly on the active samples
    acc, rgb, depth = whole['acc_map'], whole['rgb_map'], whole['depth_map']
    assert float(acc.min()) >= 0.0 and float(acc.max()) <= 1.0 + 1e-5
    assert float(rgb.min()) >= 0.0 and float(rgb.max()) <= 1.0 + 1e-5
    assert float(depth.min()) >= 0.0 and float(depth.max()) <= float(far.max()) * (1.0 + 1e-5)
    empty = ~(raw[..., :3] != 0).any(-1).any(-1)
    assert float(acc[empty].abs().max()) == 0.0 and float(rgb[empty].abs().max()) == 0.0


@pytest.mark.parametrize('pose_seed,voxel,azimuth,latent,white', [(5, 0.05, 0.9, 0, False), (9, 0.04, 2.4, 17, True), (13, 0.06, 4.0, 59, False)])
def test_render_other_poses_views_and_backgrounds(dev, pose_seed, voxel, azimuth, latent, white):
    """Different synthetic poses / volume resolutions / camera azimuths / latent codes, with and without cfg.white_bkgd, in both
    the render-only (compact compositing) and the full-contract (dense raw) mode: active set bit-exact, maps within 2e-3."""
    from animatable_nerf_b200 import synthetic
    frame = synthetic.make_frame(pose_seed=pose_seed, body_seed=1, voxel=voxel, latent_index=latent)
    K, R, T = synthetic.make_camera(frame, 112, 112, focal=120.0, azimuth=azimuth)
    ro, rd, near, far, mask = O.get_rays_within_bounds(112, 112, K, R, T, frame['wbounds'])
    assert ro.shape[0] > 2048                                        # more than one chunk, ragged tail
    batch = synthetic.make_render_batch(frame, ro, rd, near, far)
    sd = synthetic.make_state_dict(seed=pose_seed)
    ref = O.render(sd, batch, O.OracleCfg(perturb=0., white_bkgd=white), return_debug=True)
    for render_only in (True, False):
        r = _renderer(dev, sd, white_bkgd=white, b200_render_only=render_only)
        dv = r.render_device(to_device(batch, dev))
        assert int(dv['n_active'].item()) == int(ref['_debug']['pind'].sum())
        _check_maps({k: dv[k].view(ref[k].shape) for k in ('rgb_map', 'acc_map', 'depth_map')}, ref)
        if not render_only:
            na = int(dv['n_active'].item())
            assert np.array_equal(dv['active_index'][:na].cpu().numpy(), np.nonzero(ref['_debug']['pind'].numpy())[0].astype(np.int32))
            assert float((dv['pbw_all'][:na].cpu() - ref['_debug']['pbw_all']).abs().max()) <= BW_TOL
            assert float((dv['raw'].view(ref['raw'].shape).cpu() - ref['raw']).abs().max()) <= RGB_TOL


def test_render_frames_equals_per_frame_render(dev):
    """Renderer.render_frames (the evaluation loop of run.py:59-70 with the next frame's upload in flight) returns, frame by frame,
    exactly what render(to_device(batch)) returns -- different poses / latent indices per frame, pinned and pageable host batches."""
    from animatable_nerf_b200 import config, synthetic
    from animatable_nerf_b200.tpose_nerf_network import Network
    from animatable_nerf_b200.tpose_renderer import Renderer
    sd = synthetic.make_state_dict(seed=0)
    cfg = config.make_cfg(perturb=0., b200_render_only=True)
    net = Network(cfg)
    net.load_state_dict(sd)
    r = Renderer(net.to(dev).eval(), cfg)
    batches = []
    for i, (pose, lat) in enumerate(((3, 0), (7, 5), (11, 9), (5, 2), (13, 7))):      # five frames: the three rotating buffer sets are reused
        frame = synthetic.make_frame(pose_seed=pose, body_seed=1, voxel=0.05, latent_index=lat)
        K, R, T = synthetic.make_camera(frame, 96, 96, focal=100.0)
        ro, rd, near, far, _ = O.get_rays_within_bounds(96, 96, K, R, T, frame['wbounds'])
        b = synthetic.make_render_batch(frame, ro, rd, near, far)
        if i % 2 == 0:
            b = {k: (v.pin_memory() if torch.is_tensor(v) else v) for k, v in b.items()}
        batches.append(b)
    want = [r.render(r.to_device(b, dev)) for b in batches]
    got = list(r.render_frames(iter(batches), dev))
    assert len(got) == len(want)
    for g, w in zip(got, want):
        assert set(g) == set(w)
        for k in w:
            assert not g[k].is_cuda and torch.equal(g[k], w[k]), k
    assert list(r.render_frames(iter([]), dev)) == []
