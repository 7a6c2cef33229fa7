"""Marching cubes (aninerf_mesh_renderer.py:37-44).  CPU: the table-free oracle vs the generated 256-case table and
the surface properties every Lorensen-Cline mesh has; GPU: the CUDA kernel vs the oracle, index for index."""
import os
import re
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import marching_cubes as M  # noqa: E402


def _table_rows():
    txt = open(os.path.join(ROOT, 'animatable_nerf_b200', 'csrc', 'mc_tables.h')).read()
    tab = txt[txt.index('kMcTriTable'):]
    rows = [[int(v) for v in r.split(',')] for r in re.findall(r'\{([^{}]+)\}', tab)]
    assert len(rows) == 256
    return rows


def _edge_manifold(T):
    e = np.concatenate([T[:, [0, 1]], T[:, [1, 2]], T[:, [2, 0]]])
    key = np.sort(e, axis=1)
    _, cnt = np.unique(key, axis=0, return_counts=True)
    # orientation: every undirected edge is used once in each direction
    fwd = {(int(a), int(b)) for a, b in e}
    return cnt, all((b, a) in fwd for a, b in fwd)


def test_generated_table_equals_the_table_free_oracle_on_all_256_cases():
    """Each case as a single cell: the triangles the case table yields (tools/gen_mc_table.py) are the triangles the oracle traces
    from the corner values, in the same order and with the same winding; the generator's own checks (published rows) pass."""
    rows = _table_rows()
    import subprocess
    subprocess.run([sys.executable, os.path.join(ROOT, 'tools', 'gen_mc_table.py'), '--check'], check=True, capture_output=True)
    for c in range(256):
        cube = np.ones((2, 2, 2))
        for m, (dx, dy, dz) in enumerate(M._CORNER):
            if (c >> m) & 1:
                cube[dx, dy, dz] = 0.0
        V, T = M.marching_cubes(cube, 0.5)
        vid, owners = 0, {}
        for i in range(2):
            for j in range(2):
                for k in range(2):
                    for ax in range(3):
                        q = [i, j, k]
                        q[ax] += 1
                        if q[ax] < 2 and (cube[i, j, k] <= 0.5) != (cube[tuple(q)] <= 0.5):
                            owners[(i, j, k, ax)] = vid
                            vid += 1
        r = [e for e in rows[c] if e != 255]
        want = []
        for t in range(0, len(r), 3):
            want.append(tuple(owners[M._CORNER[M._EDGE_OWNER[e][0]] + (M._EDGE_OWNER[e][1],)] for e in r[t:t + 3]))
        assert [tuple(int(x) for x in t) for t in T] == want, c
        assert V.shape[0] == vid


def _blobs(n=28, seed=0):
    g = np.random.RandomState(seed)
    x, y, z = np.meshgrid(*(np.linspace(-1, 1, n),) * 3, indexing='ij')
    f = np.zeros((n, n, n))
    for _ in range(4):
        c = g.uniform(-0.4, 0.4, 3)
        f += np.exp(-((x - c[0]) ** 2 + (y - c[1]) ** 2 + (z - c[2]) ** 2) / g.uniform(0.03, 0.08))
    return (f * 60).astype(np.float32)


def test_oracle_surface_properties():
    """A sphere and a union of blobs, zero-padded like the reference's cube: closed, consistently oriented 2-manifold (every edge
    shared by exactly two triangles, once per direction), Euler characteristic 2 for the sphere, every vertex on the iso level of
    the linear interpolant, normals pointing to the <= iso side (outwards for a density)."""
    n = 24
    x, y, z = np.meshgrid(*(np.arange(n, dtype=np.float64),) * 3, indexing='ij')
    sphere = (100.0 - 12.0 * np.sqrt((x - 11.3) ** 2 + (y - 11.6) ** 2 + (z - 12.1) ** 2)).astype(np.float32)
    V, T = M.marching_cubes(np.pad(sphere.clip(min=0), 2), 50.0)
    cnt, oriented = _edge_manifold(T)
    assert (cnt == 2).all() and oriented
    assert V.shape[0] - len(cnt) + T.shape[0] == 2
    c = np.array([13.3, 13.6, 14.1])
    nrm = np.cross(V[T[:, 1]] - V[T[:, 0]], V[T[:, 2]] - V[T[:, 0]])
    assert (np.einsum('ij,ij->i', nrm, V[T].mean(1) - c) > 0).all()       # density falls outwards: normals point away from the centre
    r = np.linalg.norm(V - c, axis=1)
    assert np.abs(r - (100.0 - 50.0) / 12.0).max() < 0.08
    Vb, Tb = M.marching_cubes(np.pad(_blobs(), 3), 50.0)
    cnt, oriented = _edge_manifold(Tb)
    assert Tb.shape[0] > 500 and (cnt == 2).all() and oriented


@pytest.mark.gpu
def test_cuda_marching_cubes_equals_the_oracle():
    """aninerf_marching_cubes vs oracle/marching_cubes.py on blobs, a sphere, random noise (every ambiguous case) and degenerate
    inputs (values exactly at the iso level, all-inside / all-outside cubes): vertices bit-equal in float64, triangles identical."""
    from animatable_nerf_b200 import aninerf_mesh_renderer as R
    dev = torch.device('cuda:0')
    g = np.random.RandomState(5)
    cases = [(np.pad(_blobs(28, 1), 10), 50.0), (np.pad(_blobs(20, 2), 3), 20.0), ((g.rand(18, 13, 21) * 100).astype(np.float32), 50.0),
             (np.round(g.rand(12, 12, 12) * 4).astype(np.float32) * 25.0, 50.0), (np.zeros((6, 7, 8), np.float32), 50.0),
             (np.full((5, 5, 5), 80.0, np.float32), 50.0), ((g.rand(2, 2, 2) * 100).astype(np.float32), 50.0)]
    for cube, iso in cases:
        V, T = M.marching_cubes(cube, iso)
        gv, gt = R.marching_cubes(torch.from_numpy(cube).to(dev), iso)
        assert gv.shape[0] == V.shape[0] and gt.shape[0] == T.shape[0], (cube.shape, gv.shape, V.shape, gt.shape, T.shape)
        if V.shape[0]:
            assert np.array_equal(gv.cpu().numpy(), V)
            assert np.array_equal(gt.cpu().numpy().astype(np.int64), T)


@pytest.mark.gpu
def test_mesh_renderer_contract_and_full_size_properties():
    """aninerf_mesh_renderer.Renderer.render on a synthetic frame: the returned keys / coordinates of the reference; the mesh of the
    256^3-class padded cube is a closed oriented manifold whose vertices sit on sigma == mesh_th of the linear interpolant, and it
    equals the oracle's mesh of the same cube."""
    from animatable_nerf_b200 import aninerf_mesh_renderer as R, config, sweep, synthetic
    from animatable_nerf_b200.tpose_nerf_network import Network
    dev = torch.device('cuda:0')
    frame = synthetic.make_frame(pose_seed=2, body_seed=1, voxel=0.05, latent_index=0)
    sd = synthetic.make_state_dict(seed=0)
    # random-init densities are ~0.05: scale the density head so that the surface sigma == mesh_th exists
    sd['tpose_human.alpha_fc.weight'] = sd['tpose_human.alpha_fc.weight'] * 400.0
    sd['tpose_human.alpha_fc.bias'] = sd['tpose_human.alpha_fc.bias'] * 0 + 10.0
    cfg = config.make_cfg(perturb=0., voxel_size=[0.02, 0.02, 0.02], mesh_th=5.0)
    net = Network(cfg)
    net.load_state_dict(sd)
    net = net.to(dev).eval()
    batch = synthetic.collate_frame(frame, dev)
    pts = sweep.grid_points(frame['wbounds'], cfg.voxel_size, dev)
    batch['pts'] = pts[None]
    batch['inside'] = torch.ones(1, *pts.shape[:3], dtype=torch.uint8, device=dev)
    ret = R.Renderer(net, cfg).render(batch)
    assert set(ret) >= {'vertex', 'posed_vertex', 'triangle'}
    V, T, cube = ret['vertex'], ret['triangle'].astype(np.int64), ret['cube'].cpu().numpy()
    assert T.shape[0] > 1000
    Vo, To = M.extract_mesh(cube, 5.0, 0.02, frame['wbounds'][0])
    assert np.array_equal(To, T) and np.abs(Vo - V).max() < 1e-12
    cnt, oriented = _edge_manifold(T)
    assert (cnt == 2).all() and oriented                           # the zero padding closes the surface
    lo, hi = frame['wbounds'][0] - 0.021, frame['wbounds'][1] + 0.041
    assert (V >= lo).all() and (V <= hi).all()
