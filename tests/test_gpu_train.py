"""GPU parity of the training step (BASELINE config 4): NetworkWrapper.forward + loss.backward() on the library's kernels
vs the CPU oracle's autograd (oracle/aninerf_oracle.train_step_grads, pinned bit-equal to the reference trainer) and vs
the committed REFERENCE losses / gradients (tests/golden/train_step_small.npz)."""
import numpy as np
import pytest
import torch

from helpers import O, golden_small_case, load_golden, to_device

pytestmark = pytest.mark.gpu

LOSS_TOL = 1e-5      # absolute, fp32 losses of O(0.1)
GRAD_TOL = 2e-3      # per tensor: max |g - g_ref| / max |g_ref|  (fp32-equivalent bf16x3 products; measured ~1e-4)


@pytest.fixture(scope='module')
def dev():
    assert torch.cuda.is_available(), 'these tests need the B200'
    return torch.device('cuda:0')


def _golden_train_batch():
    g, batch, sd = golden_small_case()
    gt = load_golden('train_step_small.npz')
    tb = dict(batch)
    for k in ('ray_o', 'ray_d', 'near', 'far', 'rgb', 'mask_at_box'):
        tb[k] = torch.from_numpy(gt[k])[None]
    tb['occupancy'] = torch.ones(1, tb['near'].shape[1], dtype=torch.uint8)
    return gt, tb, sd, torch.from_numpy(gt['t_rand'])


def _wrapper(dev, sd):
    from animatable_nerf_b200 import config
    from animatable_nerf_b200.tpose_nerf_network import Network
    from animatable_nerf_b200.tpose_trainer import NetworkWrapper
    cfg = config.make_cfg(perturb=1.)
    net = Network(cfg)
    net.load_state_dict(sd)
    net = net.to(dev)
    net.train()
    return NetworkWrapper(net, cfg)


def test_train_step_matches_reference_and_oracle(dev):
    gt, tb, sd, t_rand = _golden_train_batch()
    w = _wrapper(dev, sd)
    ret, loss, stats, _ = w(to_device(tb, dev), t_rand=t_rand)
    assert loss.requires_grad
    loss.mean().backward()
    torch.nn.utils.clip_grad_value_(w.net.parameters(), 40)
    # committed REFERENCE numbers
    for k in ('bw_loss', 'img_loss', 'loss'):
        assert abs(float(stats[k]) - float(gt['stat_' + k])) <= LOSS_TOL, k
    grads = {k: p.grad.detach().cpu() for k, p in w.net.named_parameters()}
    assert len(grads) == 46
    for k in [f[5:] for f in gt.files if f.startswith('grad_')]:
        ref = torch.from_numpy(gt['grad_' + k])
        err = float((grads[k] - ref).abs().max() / ref.abs().max().clamp_min(1e-12))
        assert err <= GRAD_TOL, (k, err)
    for k in grads:
        assert abs(float(grads[k].norm()) - float(gt['gradnorm_' + k])) <= GRAD_TOL * max(float(gt['gradnorm_' + k]), 1e-8), k
    # the oracle on the same inputs: every one of the 46 tensors
    stats_o, grads_o = O.train_step_grads(sd, tb, O.OracleCfg(perturb=1.), t_rand=t_rand)
    worst = ('', 0.0)
    for k, ref in grads_o.items():
        err = float((grads[k] - ref).abs().max() / ref.abs().max().clamp_min(1e-12))
        worst = max(worst, (k, err), key=lambda kv: kv[1])
        assert err <= GRAD_TOL, (k, err)
    print('worst gradient error', worst, 'loss', float(loss), stats_o['loss'])
    # the render contract of the training mode
    ref = O.render(sd, tb, O.OracleCfg(perturb=1.), t_rand=t_rand)
    for k in ('rgb_map', 'acc_map', 'depth_map', 'raw'):
        assert float((ret[k].cpu() - ref[k]).abs().max()) <= 1e-4, k
    assert ret['pbw'].shape == ref['pbw'].shape and float((ret['pbw'].cpu() - ref['pbw']).abs().max()) <= 1e-5
    assert float((ret['tbw'].cpu() - ref['tbw']).abs().max()) <= 1e-5


def test_train_iteration_reduces_the_loss(dev):
    """A few Adam iterations (lib/train/optimizer.py: lr 5e-4) on one batch: the loss goes down and the fused render path
    sees the updated weights (the packed operand images are rebuilt after optimizer.step)."""
    from animatable_nerf_b200.tpose_trainer import train_iteration
    gt, tb, sd, t_rand = _golden_train_batch()
    w = _wrapper(dev, sd)
    opt = torch.optim.Adam(w.net.parameters(), lr=5e-4)
    b = to_device(tb, dev)
    # the fast path of train_iteration (flat buffer -> .grad) carries the same gradients as loss.backward() + clip_grad_value_
    _, loss, _, _ = w(b, t_rand=t_rand)
    loss.mean().backward()
    torch.nn.utils.clip_grad_value_(w.net.parameters(), 40)
    via_autograd = {k: p.grad.clone() for k, p in w.net.named_parameters()}
    w.net.zero_grad()
    train_iteration(w, b, torch.optim.SGD(w.net.parameters(), lr=0.0), t_rand=t_rand)
    for k, p in w.net.named_parameters():
        assert torch.equal(p.grad, via_autograd[k]), k
    losses = []
    for _ in range(5):
        _, stats = train_iteration(w, b, opt, t_rand=t_rand)
        losses.append(float(stats['loss']))
    assert losses[-1] < losses[0], losses
    w.net.eval()
    out = w.renderer.render(b)
    assert torch.isfinite(out['rgb_map']).all()


def test_animation_train_step_matches_reference_and_oracle(dev):
    """Stage 2 (aninerf_animation_trainer): two blend-weight consistency losses, gradients of novel_pose_bw.* only."""
    from animatable_nerf_b200 import config, synthetic
    from animatable_nerf_b200.tpose_nerf_network import Network
    from animatable_nerf_b200.aninerf_animation_trainer import NetworkWrapper
    _, batch, _ = golden_small_case()
    g2 = load_golden('render_small_novel_pose.npz')
    ga = load_golden('animation_train_step_small.npz')
    sd2 = synthetic.make_state_dict(seed=int(g2['sd_seed']), num_eval_frame=int(g2['num_eval_frame']))
    cfg = config.make_cfg(perturb=0., aninerf_animation=True, num_eval_frame=int(g2['num_eval_frame']))
    net = Network(cfg)
    net.load_state_dict(sd2)
    net = net.to(dev).train()
    w = NetworkWrapper(net, cfg)
    wpts, tpts = torch.from_numpy(ga['wpts']), torch.from_numpy(ga['tpts'])
    ret, loss, stats, _ = w(to_device(batch, dev), wpts=wpts, tpts=tpts)
    loss.mean().backward()
    torch.nn.utils.clip_grad_value_(net.parameters(), 40)
    for k in ('bw_loss0', 'bw_loss1', 'loss'):
        assert abs(float(stats[k]) - float(ga['stat_' + k])) <= 1e-6, (k, float(stats[k]), float(ga['stat_' + k]))
    grads = {k: p.grad.detach().cpu() for k, p in net.named_parameters() if p.grad is not None}
    assert all(k.startswith('novel_pose_bw.') for k in grads) and len(grads) == 19
    for k in [f[5:] for f in ga.files if f.startswith('grad_')]:
        ref = torch.from_numpy(ga['grad_' + k])
        err = float((grads[k] - ref).abs().max() / ref.abs().max().clamp_min(1e-12))
        assert err <= GRAD_TOL, (k, err)
    stats_o, grads_o = O.animation_train_step_grads(sd2, batch, wpts[None], tpts[None], O.OracleCfg())
    worst = ('', 0.0)
    for k, ref in grads_o.items():
        err = float((grads[k] - ref).abs().max() / ref.abs().max().clamp_min(1e-12))
        worst = max(worst, (k, err), key=lambda kv: kv[1])
        assert err <= GRAD_TOL, (k, err)
    print('stage-2 worst gradient error', worst, 'loss', float(loss), stats_o['loss'])
