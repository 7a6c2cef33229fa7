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
    # the training-step, culling and tiled-render entries validate their arguments the same way (no GPU is touched)
    import ctypes as C
    g = _lib.Gemm()
    g.n_seg, g.M, g.N = 1, 8, 8
    assert L.aninerf_gemm_x3(C.byref(g), None, 0, None) == -1                         # no output / operand pointers
    assert L.aninerf_colsum(None, 8, 8, 8, None, 0, None, 0, None) == -1
    assert L.aninerf_pe_forward(None, 8, 10, None, 64, None) == -1
    assert L.aninerf_bw_softmax_forward(None, 25, None, 8, None, None) == -1
    assert L.aninerf_inverse_lbs_backward(None, None, None, None, 8, None, 0, None) == -1
    assert L.aninerf_composite_backward(None, None, 8, 64, 0, None, None) == -1
    assert L.aninerf_img_loss(None, None, None, 8, None, None, None) == -1
    assert L.aninerf_inside_all_views(None, 8, None, None, None) == -1
    fr, pr, ro = _lib.Frame(), _lib.RenderParams(n_samples=64, chunk_rays=2048), _lib.RenderOutputs()
    pg = _lib.PeerGather()
    pg.world, pg.rank = 9, 0                                                            # more peers than one NVSwitch box has
    assert L.aninerf_render_rays_tiled(None, C.byref(fr), C.byref(pr), None, C.byref(pg), None, None, None, None, None, None, 4,
                                       C.byref(ro), None, 0, None) == -1
    assert L.aninerf_render_rays_culled(None, C.byref(fr), C.byref(pr), None, None, None, None, None, None, None, 4, C.byref(ro), None, 0,
                                        None) == -1


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


def test_ctypes_struct_layouts_match_the_header(tmp_path):
    """The ctypes mirrors in animatable_nerf_b200/_lib.py against the C compiler's view of include/aninerf_b200.h: sizes and the
    offset of every struct's last field (a mismatch would silently shift pointers across the C ABI)."""
    import ctypes as C
    import subprocess
    from animatable_nerf_b200 import _lib
    pairs = [('aninerf_camera', _lib.Camera, 'W'), ('aninerf_layer', _lib.Layer, 'relu'), ('aninerf_frame', _lib.Frame, 'bw_latent_index_dev'),
             ('aninerf_render_params', _lib.RenderParams, 'nerf_precision'), ('aninerf_render_outputs', _lib.RenderOutputs, 'chunk_offsets'),
             ('aninerf_silhouettes', _lib.Silhouettes, 'W'), ('aninerf_peer_gather', _lib.PeerGather, 'rank'),
             ('aninerf_gemm_seg', _lib.GemmSeg, 'K'), ('aninerf_gemm', _lib.Gemm, 'split_k')]
    src = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{os.path.join(ROOT, "include", "aninerf_b200.h")}"', 'int main(void) {']
    for cname, _, last in pairs:
        src.append(f'  printf("%zu %zu\\n", sizeof({cname}), offsetof({cname}, {last}));')
    src += ['  return 0;', '}']
    c_file, exe = tmp_path / 'layout.c', tmp_path / 'layout'
    c_file.write_text('\n'.join(src))
    subprocess.run(['gcc', '-std=c99', '-o', str(exe), str(c_file)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    for i, (cname, ct, last) in enumerate(pairs):
        size, off = int(out[2 * i]), int(out[2 * i + 1])
        assert C.sizeof(ct) == size, (cname, C.sizeof(ct), size)
        assert getattr(ct, last).offset == off, (cname, last)
