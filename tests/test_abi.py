"""CPU: the C-ABI shared library loads and exports every symbol include/aninerf_b200.h declares
(no compute calls: there is no GPU here)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, 'include', 'aninerf_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(aninerf_[a-z0-9_]+)\s*\(', text)))


def test_header_declares_the_path():
    names = _declared()
    for must in ('aninerf_gen_rays', 'aninerf_near_far', 'aninerf_sample_points', 'aninerf_sample_blend_weights', 'aninerf_inverse_lbs',
                 'aninerf_bw_forward', 'aninerf_nerf_forward', 'aninerf_composite', 'aninerf_render_rays', 'aninerf_query_alpha'):
        assert must in names


def test_library_exports_every_declared_symbol():
    from animatable_nerf_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH), 'build it first: python -c "import __graft_entry__ as g; g.build()"'
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in _declared():
        assert hasattr(lib, name), f'{name} is declared in include/aninerf_b200.h but not exported'
    # and the Python binding covers the same set
    assert set(_lib.PROTOTYPES) == set(_declared())


def test_version_and_error_string_without_a_gpu():
    from animatable_nerf_b200 import _lib
    L = _lib.lib()
    assert L.aninerf_version() == 1
    assert isinstance(L.aninerf_last_error(), bytes)
    assert L.aninerf_compact_workspace_bytes(1 << 20) > 0
    assert L.aninerf_render_workspace_bytes(2048, 64, 1, 1000, 1000) > 2048 * 64 * 44


def test_bad_arguments_are_rejected_not_crashed():
    from animatable_nerf_b200 import _lib
    L = _lib.lib()
    rc = L.aninerf_composite(None, None, 10, 64, 0, None, None, None, None, None, None)
    assert rc == -1 and b'invalid argument' in L.aninerf_last_error()
    rc = L.aninerf_sample_points(None, None, None, None, None, None, 4, 64, None, None, None, None)
    assert rc == -1


def test_no_cpu_fallback():
    """The product path must fail loudly off-GPU instead of computing on the host."""
    import pytest
    import torch
    from animatable_nerf_b200 import AninerfError, config
    from animatable_nerf_b200.tpose_nerf_network import Network
    from animatable_nerf_b200.tpose_renderer import Renderer
    net = Network(config.make_cfg())
    with pytest.raises(AninerfError):
        net.packed()
    batch = {'ray_o': torch.zeros(1, 4, 3), 'ray_d': torch.zeros(1, 4, 3), 'near': torch.zeros(1, 4), 'far': torch.ones(1, 4)}
    with pytest.raises(AninerfError):
        Renderer(net).render(batch)


def test_product_package_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, 'animatable_nerf_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                src = open(os.path.join(dirpath, f)).read()
                assert 'import oracle' not in src and 'from oracle' not in src, f
