"""Run by tests/test_reference_seam.py in a subprocess (the reference import mutates process-global state).

Drives the drop-in exactly as a maintainer would: the UNMODIFIED reference's `lib.config.cfg` with the four overrides of
INTEGRATION.md section 2, then the reference's own factories and checkpoint loader:
  lib/networks/make_network.py:5-9      -> animatable_nerf_b200.tpose_nerf_network.Network
  lib/networks/renderer/make_renderer.py:5-9 -> animatable_nerf_b200.tpose_renderer.Renderer
  lib/utils/net_utils.py:load_network   -> loads a checkpoint written from the REFERENCE's Network into the drop-in
CPU only: nothing here launches a kernel (construction, state_dict, checkpoint IO).
"""
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from oracle import reference_import  # noqa: E402

PKG = os.path.join(ROOT, 'animatable_nerf_b200')


def main():
    over = ['network_module', 'animatable_nerf_b200.tpose_nerf_network', 'network_path', os.path.join(PKG, 'tpose_nerf_network.py'),
            'renderer_module', 'animatable_nerf_b200.tpose_renderer', 'renderer_path', os.path.join(PKG, 'tpose_renderer.py'),
            'trainer_module', 'animatable_nerf_b200.tpose_trainer', 'trainer_path', os.path.join(PKG, 'tpose_trainer.py')]
    ref = reference_import.load('configs/aninerf_313.yaml', overrides=over)
    cfg = ref.cfg
    assert cfg.network_path.endswith('animatable_nerf_b200/tpose_nerf_network.py')
    os.chdir(reference_import.REF)
    from lib.networks import make_network            # lib/networks/__init__.py re-exports the factory function
    from lib.networks.renderer import make_renderer
    from lib.utils import net_utils
    import animatable_nerf_b200.tpose_nerf_network as ours_net
    import animatable_nerf_b200.tpose_renderer as ours_ren
    torch.manual_seed(0)
    net = make_network(cfg)
    assert type(net).__module__ == 'animatable_nerf_b200.tpose_nerf_network' and type(net).__name__ == 'Network'
    ren = make_renderer(cfg, net)
    assert type(ren).__module__ == 'animatable_nerf_b200.tpose_renderer' and type(ren).__name__ == 'Renderer'
    assert ren.net is net and ren.cfg is cfg                    # the process-global reference cfg drives the drop-in
    assert hasattr(ours_net, 'Network') and hasattr(ours_ren, 'Renderer')
    # the reference's own Network, loaded by file, same seed: keys, shapes AND order (and, with the same seed, values)
    import importlib
    ref_mod = importlib.import_module('lib.networks.bw_deform.tpose_nerf_network')      # the reference's own module
    assert ref_mod.__file__.startswith(reference_import.REF)
    torch.manual_seed(0)
    ref_net = ref_mod.Network()
    sd_ref, sd_ours = ref_net.state_dict(), net.state_dict()
    assert list(sd_ref.keys()) == list(sd_ours.keys()), 'state_dict keys / order differ'
    assert [tuple(v.shape) for v in sd_ref.values()] == [tuple(v.shape) for v in sd_ours.values()]
    n_params = sum(v.numel() for v in sd_ours.values())
    assert len(sd_ours) == 46 and n_params == 1274652, (len(sd_ours), n_params)
    assert all(torch.equal(sd_ref[k], sd_ours[k]) for k in sd_ref), 'seeded default init differs (construction order)'
    # a checkpoint written from the REFERENCE network, read by the REFERENCE's load_network into the drop-in
    with tempfile.TemporaryDirectory() as d:
        for p in ref_net.parameters():
            torch.nn.init.normal_(p, std=0.02)
        torch.save({'net': ref_net.state_dict(), 'optim': {}, 'scheduler': {}, 'recorder': {}, 'epoch': 41}, os.path.join(d, 'latest.pth'))
        assert net_utils.load_network(net, d, resume=True) == 42
        assert all(torch.equal(a, b) for a, b in zip(ref_net.state_dict().values(), net.state_dict().values()))
        # ... and the drop-in's own checkpoint module reads the same file
        from animatable_nerf_b200 import checkpoint
        net2 = make_network(cfg)
        assert checkpoint.load_network(net2, d) == 42
        assert all(torch.equal(a, b) for a, b in zip(ref_net.state_dict().values(), net2.state_dict().values()))
    # the trainer seam: tpose_trainer.NetworkWrapper(net) as lib/train/trainers/make_trainer.py builds it
    tr_mod = reference_import._load_source(cfg.trainer_module, cfg.trainer_path)
    w = tr_mod.NetworkWrapper(net)
    assert w.net is net and hasattr(w, 'renderer')
    # no CPU fallback: the drop-in refuses host tensors instead of computing on the CPU
    from animatable_nerf_b200 import _lib
    try:
        ren.render({'ray_o': torch.zeros(1, 4, 3), 'ray_d': torch.zeros(1, 4, 3), 'near': torch.zeros(1, 4), 'far': torch.ones(1, 4)})
    except _lib.AninerfError:
        pass
    else:
        raise AssertionError('Renderer.render accepted CPU tensors')
    print('REF_SEAM_OK', len(sd_ours), n_params)


if __name__ == '__main__':
    main()
