"""CPU oracle for the mesh extraction step of `aninerf_mesh_renderer.Renderer.render`
(lib/networks/renderer/aninerf_mesh_renderer.py:37-44): `mcubes.marching_cubes(np.pad(cube, 10), cfg.mesh_th)`.

TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED: the arithmetic lives in PyMCubes 0.1.0 (requirements.txt:11), a third-party
dependency that is neither vendored in /root/reference nor installed in this image, and the reference has no test or fixture
for it.  This file restates the published algorithm PyMCubes implements (Lorensen & Cline marching cubes with P. Bourke's
corner / edge numbering): corner bit set when value <= iso; one vertex per cube edge whose end points differ, at
`p1 + (iso - f1) (p2 - p1) / (f2 - f1)` in float64 index coordinates; per cell the crossed edges are joined face by face
(every set corner of an ambiguous face is cut off on its own), the segments close into loops, every loop is fan-triangulated.
It is written WITHOUT the 256-case table -- per active cell it traces the loops from the corner values -- so that it checks the
table-driven CUDA kernel (and tools/gen_mc_table.py) rather than sharing their data.

Conventions shared with the CUDA kernel (they make results directly comparable, index for index):
  vertices  ordered by owning grid point (x-major linear index), then by axis (x, y, z edge leaving that point);
  triangles ordered by cell (x-major), inside a cell by loop (ascending smallest edge id), fan from the loop's smallest edge id,
            wound so that the normal points to the set (<= iso) side.
"""
from __future__ import annotations

import numpy as np

_CORNER = ((0, 0, 0), (1, 0, 0), (1, 1, 0), (0, 1, 0), (0, 0, 1), (1, 0, 1), (1, 1, 1), (0, 1, 1))
_EDGE = ((0, 1), (1, 2), (2, 3), (3, 0), (4, 5), (5, 6), (6, 7), (7, 4), (0, 4), (1, 5), (2, 6), (3, 7))
# edge -> (corner that owns it, axis): the edge leaves the owner in +axis direction
_EDGE_OWNER = ((0, 0), (1, 1), (3, 0), (0, 1), (4, 0), (5, 1), (7, 0), (4, 1), (0, 2), (1, 2), (2, 2), (3, 2))
_FACE = ((0, 1, 2, 3), (4, 5, 6, 7), (0, 1, 5, 4), (1, 2, 6, 5), (2, 3, 7, 6), (3, 0, 4, 7))


def _edge_id(a, b):
    for e, (p, q) in enumerate(_EDGE):
        if (p, q) == (a, b) or (p, q) == (b, a):
            return e
    raise KeyError((a, b))


def _cell_loops(bits):
    """bits[m] = corner m is set.  Returns loops of edge ids, each starting at its smallest edge id and running so that, seen from
    outside the cube, the set corners lie to the left of every segment."""
    nxt = {}                                              # directed successor per edge id
    for f in _FACE:
        pts = np.array([_CORNER[v] for v in f], float)
        normal = np.sign(pts.mean(0) - 0.5)
        fe = [_edge_id(f[i], f[(i + 1) % 4]) for i in range(4)]
        cross = [i for i in range(4) if bits[f[i]] != bits[f[(i + 1) % 4]]]
        segs = []
        if len(cross) == 2:
            segs.append((fe[cross[0]], fe[cross[1]], next(f[i] for i in range(4) if bits[f[i]])))
        elif len(cross) == 4:
            segs += [(fe[(i - 1) % 4], fe[i], f[i]) for i in range(4) if bits[f[i]]]
        for a, b, corner in segs:
            pa = (np.array(_CORNER[_EDGE[a][0]], float) + np.array(_CORNER[_EDGE[a][1]], float)) / 2
            pb = (np.array(_CORNER[_EDGE[b][0]], float) + np.array(_CORNER[_EDGE[b][1]], float)) / 2
            left = np.cross(normal, pb - pa)
            if np.dot(left, np.array(_CORNER[corner], float) - (pa + pb) / 2) > 0:
                nxt[a] = b
            else:
                nxt[b] = a
    loops, seen = [], set()
    for start in sorted(nxt):
        if start in seen:
            continue
        loop, cur = [], start
        while cur not in seen:
            seen.add(cur)
            loop.append(cur)
            cur = nxt[cur]
        assert cur == start, 'open loop'
        loops.append(loop)
    return loops


def marching_cubes(cube, iso):
    """cube (X,Y,Z) -> vertices (V,3) float64 (index coordinates), triangles (T,3) int64."""
    f = np.asarray(cube, dtype=np.float64)
    X, Y, Z = f.shape
    inside = f <= iso
    # vertices: per grid point, the x / y / z edge leaving it
    vid = -np.ones((X, Y, Z, 3), dtype=np.int64)
    verts = []
    cross = [np.zeros((X, Y, Z), bool) for _ in range(3)]
    cross[0][:-1] = inside[:-1] != inside[1:]
    cross[1][:, :-1] = inside[:, :-1] != inside[:, 1:]
    cross[2][:, :, :-1] = inside[:, :, :-1] != inside[:, :, 1:]
    any_cross = cross[0] | cross[1] | cross[2]
    for i, j, k in zip(*np.nonzero(any_cross)):
        for ax in range(3):
            if cross[ax][i, j, k]:
                q = [i, j, k]
                q[ax] += 1
                f1, f2 = f[i, j, k], f[tuple(q)]
                p = np.array([i, j, k], dtype=np.float64)
                p[ax] = p[ax] + (iso - f1) * 1.0 / (f2 - f1)
                vid[i, j, k, ax] = len(verts)
                verts.append(p)
    # triangles: per active cell
    c = inside[:-1, :-1, :-1].astype(np.int32) * 0
    for m, (dx, dy, dz) in enumerate(_CORNER):
        c |= inside[dx:X - 1 + dx, dy:Y - 1 + dy, dz:Z - 1 + dz].astype(np.int32) << m
    tris = []
    for i, j, k in zip(*np.nonzero((c != 0) & (c != 255))):
        bits = [(int(c[i, j, k]) >> m) & 1 for m in range(8)]
        for loop in _cell_loops(bits):
            ids = []
            for e in loop:
                owner, ax = _EDGE_OWNER[e]
                dx, dy, dz = _CORNER[owner]
                v = vid[i + dx, j + dy, k + dz, ax]
                assert v >= 0
                ids.append(v)
            for t in range(1, len(ids) - 1):
                tris.append((ids[0], ids[t], ids[t + 1]))
    V = np.array(verts, dtype=np.float64).reshape(-1, 3)
    T = np.array(tris, dtype=np.int64).reshape(-1, 3)
    return V, T


def extract_mesh(cube, iso, voxel_size, wbounds_min, pad=10):
    """aninerf_mesh_renderer.py:39-43: pad the sigma cube by 10 with zeros, marching cubes at cfg.mesh_th, vertices to world
    coordinates `(v - 10) * voxel_size[0] + wbounds[0, 0]`."""
    V, T = marching_cubes(np.pad(np.asarray(cube, dtype=np.float64), pad, mode='constant'), iso)
    return (V - pad) * voxel_size + np.asarray(wbounds_min, dtype=np.float64), T
