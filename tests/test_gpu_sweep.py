"""GPU parity of the frame-level drivers (BASELINE config 5): novel-view sweep and density grid vs the oracle."""
import numpy as np
import pytest
import torch

from helpers import O, golden_small_case, load_golden, to_device
from animatable_nerf_b200 import host_geometry, synthetic

pytestmark = pytest.mark.gpu
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


def _net(dev, sd):
    from animatable_nerf_b200 import config
    from animatable_nerf_b200.tpose_nerf_network import Network
    cfg = config.make_cfg(perturb=0., b200_render_only=True)
    net = Network(cfg)
    net.load_state_dict(sd)
    return net.to(dev).eval(), cfg


def test_view_sweep_matches_oracle_per_view(dev):
    """3 views of a circular path at 80x80: the rays come from the on-device front end (bit-exact vs the oracle's numpy
    stage 1), the image is the oracle's render of the same rays scattered through mask_at_box."""
    from animatable_nerf_b200 import sweep
    from animatable_nerf_b200.tpose_renderer import Renderer
    g, batch, sd = golden_small_case()
    frame = {k: g['frame_' + k] for k in synthetic.FRAME_KEYS}
    rig = synthetic.make_camera_rig(frame, n_views=5)
    path = host_geometry.circular_camera_path(list(rig), 3)
    H = W = 80
    K = np.array([[90., 0, 40.], [0, 90., 40.], [0, 0, 1.]])
    net, cfg = _net(dev, sd)
    fb = {k: v for k, v in to_device(batch, dev).items() if k not in ('ray_o', 'ray_d', 'near', 'far', 'occupancy')}
    out = sweep.render_views(Renderer(net, cfg), fb, K, path, H, W)
    assert sorted(out) == [0, 1, 2]
    hit_any = 0
    for v, (rgb, acc, depth) in out.items():
        RT = path[v]
        ro, rd, near, far, mask = O.get_rays_within_bounds(H, W, K, RT[:3, :3], RT[:3, 3:], frame['wbounds'])
        hit_any += int(mask.sum())
        img = np.zeros((H * W, 3), np.float32)
        if mask.sum():
            ref = O.render(sd, synthetic.make_render_batch(frame, ro, rd, near, far), O.OracleCfg(perturb=0.))
            img[mask.reshape(-1)] = ref['rgb_map'][0].numpy()
        assert np.abs(rgb.cpu().numpy().reshape(-1, 3) - img).max() <= RGB_TOL, v
        assert float(acc.cpu().view(-1)[~torch.from_numpy(mask.reshape(-1))].abs().max()) == 0.0
    assert hit_any > 0
    stack = sweep.gather_views(out, 3, H, W, 0, 1, dev)
    assert stack.shape == (3, H, W, 3) and torch.equal(stack[1], out[1][0])


def test_density_grid_matches_oracle(dev):
    """A coarse voxel grid over wbounds with an `inside` mask, chunks of 512 points: cube == oracle calculate_alpha
    per chunk (norm_th 0.1 mask bit-exact, per-chunk argmin forcing)."""
    from animatable_nerf_b200 import sweep
    g, batch, sd = golden_small_case()
    net, cfg = _net(dev, sd)
    pts = sweep.grid_points(batch['wbounds'][0].numpy(), [0.09, 0.09, 0.09], dev)
    gen = torch.Generator().manual_seed(3)
    inside = (torch.rand(pts.shape[:-1], generator=gen) < 0.8).to(dev)
    chunk = 512
    cube = sweep.query_density_grid(net, to_device(batch, dev), pts, inside, chunk=chunk)
    assert cube.shape == pts.shape[:-1]
    flat = pts.reshape(-1, 3).cpu()[inside.reshape(-1).cpu()]
    assert flat.shape[0] > chunk                                                # at least two chunks
    ref = torch.cat([O.calculate_alpha(sd, flat[i:i + chunk], batch, O.OracleCfg()) for i in range(0, flat.shape[0], chunk)])
    got = cube.reshape(-1)[inside.reshape(-1)].cpu()
    assert np.array_equal((got != 0).numpy(), (ref != 0).numpy())
    assert float((got - ref).abs().max()) <= 5e-3
    assert float(cube.reshape(-1)[~inside.reshape(-1)].abs().max()) == 0.0
