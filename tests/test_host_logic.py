"""CPU: host-side logic of the drop-in -- weight folding, row selection, ray-tile sharding, batch schema."""
import numpy as np
import torch

from helpers import O, golden_small_case
from animatable_nerf_b200 import config, ray_tiles, synthetic, weights
from animatable_nerf_b200.tpose_nerf_network import Network


def _run_layers(layers, x, final_relu):
    h = x
    for i, (W, b) in enumerate(layers):
        h = h @ W.T + b
        if i < len(layers) - 1 or final_relu:
            h = np.maximum(h, 0)
    return h


def test_network_mirror_has_the_reference_checkpoint_layout():
    sd = synthetic.make_state_dict(seed=0, num_train_frame=60, num_eval_frame=7)
    net = Network(config.make_cfg(aninerf_animation=True, num_eval_frame=7))
    assert sum(p.numel() for p in Network(config.make_cfg()).parameters()) == 1274652
    missing, unexpected = net.load_state_dict(sd, strict=True)
    assert not missing and not unexpected
    assert {k: tuple(v.shape) for k, v in net.state_dict().items()} == {k: tuple(v.shape) for k, v in sd.items()}


def test_blend_weight_field_folding_matches_oracle():
    """latent folded into per-index bias tables == the reference's concat-then-conv (fp64 vs fp32 noise only)"""
    sd = synthetic.make_state_dict(seed=3)
    g = torch.Generator().manual_seed(0)
    pts = torch.rand(1, 257, 3, generator=g) * 2 - 1
    smpl = torch.softmax(torch.randn(1, 24, 257, generator=g), dim=1)
    for idx in (0, 5, 60):
        ref = O.neural_blend_weights(sd, pts, smpl, torch.tensor([idx]))
        layers = weights.fold_bw_field(sd)
        assert [w.shape for w, _ in layers] == [(256, 63)] + [(256, 256)] * 4 + [(256, 319)] + [(256, 256)] * 2 + [(24, 256)]
        pe = O.positional_encoding(pts, 10)[0].double().numpy()
        h = pe
        for i, (W, b) in enumerate(layers[:8]):
            inp = np.concatenate([pe, h], axis=1) if i == 5 else h
            h = np.maximum(inp @ W.T + b[min(idx, len(b) - 1)], 0)
        logits = h @ layers[8][0].T + layers[8][1][0] + np.log(smpl[0].t().double().numpy() + 1e-9)
        bw = torch.softmax(torch.from_numpy(logits), dim=1).t()[None]
        assert float((bw - ref.double()).abs().max()) < 5e-6


def test_nerf_field_folding_matches_oracle():
    """feature_fc o latent_fc o view_fc collapsed into one (256+27)->128 layer with per-frame bias"""
    sd = synthetic.make_state_dict(seed=4)
    g = torch.Generator().manual_seed(1)
    pts = torch.rand(1, 129, 3, generator=g) * 2 - 1
    vd = torch.nn.functional.normalize(torch.randn(1, 129, 3, generator=g), dim=2)
    for idx in (0, 17):
        alpha, rgb = O.nerf_alpha_rgb(sd, pts, vd, torch.tensor([idx]))
        layers, (aw, ab, rw, rb) = weights.fold_nerf_field(sd)
        assert layers[8][0].shape == (128, 283) and layers[8][1].shape == (60, 128)
        pe = O.positional_encoding(pts, 10)[0].double().numpy()
        pv = O.positional_encoding(vd, 4)[0].double().numpy()
        h = pe
        for i, (W, b) in enumerate(layers[:8]):
            inp = np.concatenate([pe, h], axis=1) if i == 5 else h
            h = np.maximum(inp @ W.T + b[0], 0)
        a = h @ aw.T + ab
        v = np.maximum(np.concatenate([h, pv], axis=1) @ layers[8][0].T + layers[8][1][idx], 0)
        c = v @ rw.T + rb
        assert np.abs(a[:, 0] - alpha[0, 0].double().numpy()).max() < 1e-5
        assert np.abs(c.T - rgb[0].double().numpy()).max() < 1e-5


def test_ray_tiles_partition_is_a_permutation_aligned_to_chunks():
    for n_rays in (1, 2047, 2048, 2049, 242907):
        for world in (1, 2, 4, 8):
            parts = [ray_tiles.shard_indices(n_rays, r, world) for r in range(world)]
            allidx = torch.cat(parts)
            assert allidx.numel() == n_rays and torch.equal(torch.sort(allidx)[0], torch.arange(n_rays))
            assert ray_tiles.shard_sizes(n_rays, world) == [p.numel() for p in parts]
            for p in parts:                      # every piece is a run of whole reference chunks
                if p.numel():
                    starts = p[torch.cat([torch.tensor([True]), p[1:] != p[:-1] + 1])]
                    assert bool((starts % ray_tiles.CHUNK == 0).all())


def test_sharded_oracle_render_equals_unsharded():
    """Chunk-aligned shards reproduce the 1-rank result exactly (per-chunk argmin forcing is shard-local)."""
    _, batch, sd = golden_small_case()
    ref = O.render(sd, batch, O.OracleCfg(perturb=0.))
    n = batch['ray_o'].shape[1]
    out = torch.empty(n, 3)
    for r in range(2):
        sb = ray_tiles.shard_batch(batch, r, 2)
        part = O.render(sd, sb, O.OracleCfg(perturb=0.))
        out[ray_tiles.shard_indices(n, r, 2)] = part['rgb_map'][0]
    assert torch.equal(out, ref['rgb_map'][0])


def test_synthetic_batch_schema():
    frame = synthetic.make_frame(voxel=0.1)
    K, R, T = synthetic.make_camera(frame, 32, 32, focal=33.0)
    ray_o, ray_d, near, far, mask = O.get_rays_within_bounds(32, 32, K, R, T, frame['wbounds'])
    b = synthetic.make_render_batch(frame, ray_o, ray_d, near, far)
    n = ray_o.shape[0]
    assert b['ray_o'].shape == (1, n, 3) and b['near'].shape == (1, n) and b['occupancy'].dtype == torch.uint8
    assert b['A'].shape == (1, 24, 4, 4) and b['pbw'].shape[-1] == 25 and b['tbw'].shape[-1] == 25
    assert b['pbounds'].shape == (1, 2, 3) and b['R'].shape == (1, 3, 3) and b['Th'].shape == (1, 1, 3)
    assert b['latent_index'].dtype == torch.int64
    # skinning weights in the volumes are convex combinations
    assert torch.allclose(b['pbw'][..., :24].sum(-1), torch.ones_like(b['pbw'][..., 0]), atol=1e-5)


def test_sweep_and_grid_sharding_cover_every_unit_once():
    from animatable_nerf_b200 import sweep
    for n_views, world in ((64, 8), (50, 4), (3, 8), (1, 1)):
        got = sorted(v for r in range(world) for v in sweep.views_of_rank(n_views, r, world))
        assert got == list(range(n_views))
    for n_pts, world in ((256 ** 3, 8), (131072 * 3 + 17, 2), (5, 4)):
        cs = [sweep.chunks_of_rank(n_pts, r, world) for r in range(world)]
        flat = sorted(c for x in cs for c in x)
        assert flat == list(range((n_pts + sweep.GRID_CHUNK - 1) // sweep.GRID_CHUNK))
        assert max(len(x) for x in cs) - min(len(x) for x in cs) <= 1


def test_circular_camera_path_is_a_closed_orbit_of_rigid_transforms():
    import numpy as np
    from animatable_nerf_b200 import host_geometry, synthetic
    frame = synthetic.make_frame(voxel=0.1)
    _, _, RT = synthetic.make_silhouettes(frame, n_views=6, H=64, W=64, focal=66.0)
    path = host_geometry.circular_camera_path(list(RT.astype(np.float64)), 16)
    assert len(path) == 16
    centre = np.mean([np.linalg.inv(m)[:3, 3] for m in path], axis=0)
    for m in path:
        assert m.shape == (4, 4)
        assert np.allclose(m[:3, :3] @ m[:3, :3].T, np.eye(3), atol=1e-9) and abs(np.linalg.det(m[:3, :3]) - 1) < 1e-9
        pos = np.linalg.inv(m)[:3, 3]
        look = np.linalg.inv(m)[:3, 2]                    # camera z axis in the world
        assert np.dot(look, centre - pos) > 0               # every camera looks towards the orbit's inside


def test_checkpoint_layout_round_trip(tmp_path):
    """The reference's checkpoint dict ({net, optim, scheduler, recorder, epoch}, <epoch>.pth / latest.pth,
    lib/utils/net_utils.py:288-396) written and read back; `only=` keeps a sub-field as load_network does."""
    import torch
    from animatable_nerf_b200 import checkpoint, config, synthetic
    from animatable_nerf_b200.tpose_nerf_network import Network
    cfg = config.make_cfg(aninerf_animation=True, num_eval_frame=5)
    net = Network(cfg)
    net.load_state_dict(synthetic.make_state_dict(seed=3, num_eval_frame=5))
    opt = torch.optim.Adam(net.parameters(), lr=5e-4)
    sched = torch.optim.lr_scheduler.ExponentialLR(opt, 0.9)
    d = str(tmp_path / 'model')
    for ep in (3, 7):
        checkpoint.save_model(net, opt, sched, None, d, ep)
    checkpoint.save_model(net, opt, sched, None, d, 9, last=True)
    state = torch.load(d + '/7.pth', map_location='cpu')
    assert set(state) == {'net', 'optim', 'scheduler', 'recorder', 'epoch'} and state['epoch'] == 7
    assert set(state['net']) == set(net.state_dict())
    net2 = Network(cfg)
    opt2 = torch.optim.Adam(net2.parameters(), lr=1e-3)
    assert checkpoint.load_model(net2, opt2, torch.optim.lr_scheduler.ExponentialLR(opt2, 0.9), None, d) == 10      # latest.pth wins
    assert all(torch.equal(a, b) for a, b in zip(net.state_dict().values(), net2.state_dict().values()))
    net3 = Network(cfg)
    before = net3.tpose_human.alpha_fc.weight.clone()
    assert checkpoint.load_network(net3, d, epoch=3, only=['novel_pose_bw']) == 4
    assert torch.equal(net3.novel_pose_bw.bw_fc.weight, net.novel_pose_bw.bw_fc.weight) and torch.equal(net3.tpose_human.alpha_fc.weight, before)
    assert checkpoint.load_network(net3, str(tmp_path / 'missing')) == 0


def test_load_frame_from_sequence_files(tmp_path):
    """A processed-sequence directory (vertices / params / bweights / tbw / tvertices / joints / parents .npy,
    lib/datasets/tpose_dataset.py:125-161) read back into the frame dict the renderer consumes."""
    import numpy as np
    from animatable_nerf_b200 import checkpoint, host_geometry, synthetic
    frame = synthetic.make_frame(voxel=0.1)
    tverts, w, J = synthetic.make_body(1)
    root, lbs = tmp_path / 'seq', tmp_path / 'seq' / 'lbs'
    for p in (root / 'new_vertices', root / 'new_params', lbs / 'bweights'):
        p.mkdir(parents=True)
    rng = np.random.RandomState(2)
    poses = rng.normal(0, 0.2, (24, 3))
    poses[0] = 0
    Rh = rng.normal(0, 0.3, 3)
    np.save(root / 'new_vertices' / '4.npy', frame['wverts'])
    np.save(root / 'new_params' / '4.npy', {'Rh': Rh[None], 'Th': frame['Th'], 'poses': poses.reshape(1, 72), 'shapes': np.zeros((1, 10))},
            allow_pickle=True)
    np.save(lbs / 'bweights' / '4.npy', frame['pbw'])
    np.save(lbs / 'tbw.npy', frame['tbw'])
    np.save(lbs / 'tvertices.npy', tverts)
    np.save(lbs / 'joints.npy', J)
    np.save(lbs / 'parents.npy', host_geometry.SMPL_PARENTS)
    got = checkpoint.load_frame(str(root), str(lbs), 4, latent_index=2)
    assert set(synthetic.FRAME_KEYS) <= set(got)
    for k in ('A', 'R', 'pbw', 'tbw', 'wbounds', 'tbounds'):
        assert np.allclose(got[k], frame[k], atol=1e-6), k
    assert np.allclose(got['pbounds'], frame['pbounds'], atol=1e-5)
    assert got['Th'].shape == (1, 3) and int(got['latent_index']) == 2


def test_peer_scatter_row_formula_matches_the_shard_order():
    """composite_kernel's fused gather stores local ray r of rank k at frame row ((r // 2048) * world + k) * 2048 + r % 2048
    (csrc/geometry.cu, PeerScatter): exactly the frame index ray_tiles.shard_indices assigns to that local ray."""
    from animatable_nerf_b200 import ray_tiles
    for n_rays, world in ((242907, 8), (5000, 2), (2048 * 3 + 1, 4), (100, 8)):
        seen = []
        for rank in range(world):
            idx = ray_tiles.shard_indices(n_rays, rank, world)
            r = torch.arange(idx.numel())
            lc = r // ray_tiles.CHUNK
            row = (lc * world + rank) * ray_tiles.CHUNK + (r - lc * ray_tiles.CHUNK)
            assert torch.equal(row, idx), (n_rays, world, rank)
            seen.append(idx)
        assert torch.equal(torch.sort(torch.cat(seen))[0], torch.arange(n_rays))
