"""Generate the marching-cubes case tables used by csrc/marching_cubes.cu  ->  animatable_nerf_b200/csrc/mc_tables.h

The reference extracts its meshes with PyMCubes 0.1.0 (`mcubes.marching_cubes(cube, cfg.mesh_th)`,
lib/networks/renderer/aninerf_mesh_renderer.py:40), a third-party dependency that is absent from /root/reference and from this
image.  PyMCubes follows the classic Lorensen-Cline scheme with the corner / edge numbering and the 256-row triangle table that
P. Bourke published ("Polygonising a scalar field"): a corner's bit is set when its value is <= the iso value, vertices sit on
the cube edges whose end points differ, linearly interpolated.  The 4096-entry table itself is not reproduced from memory here;
it is DERIVED from the rule it implements:

  * per cube face, the crossed edges are joined pairwise; on an ambiguous face (two diagonal corners set) every SET corner is cut
    off on its own (the resolution the published table uses: e.g. row 5 = {0,8,3, 1,2,10}, row 250 = one 4-triangle patch);
  * the segments close into loops around the cube; every loop is fan-triangulated;
  * triangles are oriented so that the normal points to the set (<= iso) side, as in the published rows (row 1 = {0,8,3}).

What this guarantees: the same vertex set as any Lorensen-Cline implementation (a vertex per sign-changing edge), the same
triangle count per case as the published table, a crack-free surface.  What it does not: the same choice of interior diagonals
inside a 4..7-gon, nor the same triangle order inside a cell -- "parity unpinned" for those.
`python tools/gen_mc_table.py --check` asserts the rows of the published table that are quoted in the text above.
"""
import itertools
import os
import sys

import numpy as np

# corner m -> (dx, dy, dz);  Bourke / PyMCubes numbering: v0..v3 the z-low face counter-clockwise, v4..v7 above them
CORNERS = [(0, 0, 0), (1, 0, 0), (1, 1, 0), (0, 1, 0), (0, 0, 1), (1, 0, 1), (1, 1, 1), (0, 1, 1)]
# edge e -> (corner a, corner b)
EDGES = [(0, 1), (1, 2), (2, 3), (3, 0), (4, 5), (5, 6), (6, 7), (7, 4), (0, 4), (1, 5), (2, 6), (3, 7)]
# the six faces as corner cycles
FACES = [(0, 1, 2, 3), (4, 5, 6, 7), (0, 1, 5, 4), (1, 2, 6, 5), (2, 3, 7, 6), (3, 0, 4, 7)]
EDGE_OF = {frozenset(e): i for i, e in enumerate(EDGES)}


def edge_mask(c):
    m = 0
    for e, (a, b) in enumerate(EDGES):
        if ((c >> a) & 1) != ((c >> b) & 1):
            m |= 1 << e
    return m


FACE_NORMALS = []
for f in FACES:
    pts = np.array([CORNERS[v] for v in f], float)
    FACE_NORMALS.append(np.sign(pts.mean(0) - 0.5))           # outward normal of the unit cube's face


def loops_of(c):
    """Closed loops of cube edges for cube index c; each loop is a list of edge ids ordered so that, seen from OUTSIDE the cube,
    the set corners lie to the left of every segment (=> the triangles' normals point to the set side, as in the published rows)."""
    seg = {}                                          # frozenset(edge a, edge b) -> (face index, a set corner on the segment's set side)
    adj = {}
    for fi, f in enumerate(FACES):
        bits = [(c >> v) & 1 for v in f]
        fe = [EDGE_OF[frozenset((f[i], f[(i + 1) % 4]))] for i in range(4)]          # edge i joins corner i and i+1 of the face
        crossed = [i for i in range(4) if bits[i] != bits[(i + 1) % 4]]
        pairs = []
        if len(crossed) == 2:
            pairs.append((fe[crossed[0]], fe[crossed[1]], next(f[i] for i in range(4) if bits[i])))
        elif len(crossed) == 4:
            for i in range(4):                       # cut off every SET corner: join the two face edges incident to it
                if bits[i]:
                    pairs.append((fe[(i - 1) % 4], fe[i], f[i]))
        for a, b, corner in pairs:
            seg[frozenset((a, b))] = (fi, corner)
            adj.setdefault(a, []).append(b)
            adj.setdefault(b, []).append(a)
    assert all(len(v) == 2 for v in adj.values()), (c, adj)
    loops, seen = [], set()
    for start in sorted(adj):
        if start in seen:
            continue
        loop, prev, cur = [start], None, start
        seen.add(start)
        while True:
            nxt = [n for n in adj[cur] if n != prev] or adj[cur]
            n = nxt[0]
            if n == start:
                break
            loop.append(n)
            seen.add(n)
            prev, cur = cur, n
        # orientation: per segment a -> b on a face with outward normal nf, the set corner must lie on the side nf x (b - a)
        signs = []
        for i in range(len(loop)):
            a, b = loop[i], loop[(i + 1) % len(loop)]
            fi, corner = seg[frozenset((a, b))]
            pa, pb = mid(a), mid(b)
            left = np.cross(FACE_NORMALS[fi], pb - pa)
            signs.append(np.dot(left, np.array(CORNERS[corner], float) - (pa + pb) / 2))
        assert all(x > 1e-9 for x in signs) or all(x < -1e-9 for x in signs), (c, loop, signs)
        loop = loop if signs[0] > 0 else loop[::-1]
        i0 = loop.index(min(loop))                     # the fan's apex: the loop's smallest edge id
        loops.append(loop[i0:] + loop[:i0])
    return loops


def mid(e):
    a, b = EDGES[e]
    return (np.array(CORNERS[a], float) + np.array(CORNERS[b], float)) / 2


def tri_rows():
    rows = []
    for c in range(256):
        tris = []
        if c not in (0, 255):
            for lp in loops_of(c):
                for i in range(1, len(lp) - 1):
                    tris += [lp[0], lp[i], lp[i + 1]]
        assert len(tris) <= 15
        rows.append(tris)
    return rows


def same_triangles(row, want):
    def canon(t):
        i = t.index(min(t))
        return tuple(t[i:] + t[:i])            # rotation only: orientation must match
    a = sorted(canon(list(row[i:i + 3])) for i in range(0, len(row), 3))
    b = sorted(canon(list(want[i:i + 3])) for i in range(0, len(want), 3))
    return a == b


def check(rows):
    # edge table values of the published table (first sixteen rows) and the complement symmetry
    pub = [0x0, 0x109, 0x203, 0x30a, 0x406, 0x50f, 0x605, 0x70c, 0x80c, 0x905, 0xa0f, 0xb06, 0xc0a, 0xd03, 0xe09, 0xf00]
    assert [edge_mask(c) for c in range(16)] == pub
    assert all(edge_mask(c) == edge_mask(255 - c) for c in range(256))
    # published triangle rows quoted in the docstring (single-corner cases and the two ambiguous-face examples)
    assert same_triangles(rows[1], [0, 8, 3])
    assert same_triangles(rows[2], [0, 1, 9])
    assert same_triangles(rows[4], [1, 2, 10])
    assert same_triangles(rows[8], [3, 11, 2])
    assert same_triangles(rows[5], [0, 8, 3, 1, 2, 10])
    assert same_triangles(rows[10], [1, 9, 0, 2, 3, 11])
    assert same_triangles(rows[254], [0, 3, 8])
    assert len(rows[250]) == 12 and len(rows[3]) == 6 and len(rows[15]) == 6
    # every case: each crossed edge is used, the triangle edges pair up (closed inside the cell except along the cube faces)
    for c in range(256):
        used = set(rows[c])
        assert used == {e for e in range(12) if edge_mask(c) >> e & 1}, c
    print('mc tables: checks passed;', sum(len(r) // 3 for r in rows), 'triangles over the 256 cases')


def emit(rows, path):
    with open(path, 'w') as f:
        f.write('// GENERATED by tools/gen_mc_table.py -- do not edit.  Marching-cubes case tables (Bourke / PyMCubes corner and edge\n'
                '// numbering; a corner bit is set when value <= iso; triangles oriented towards the set side).\n#pragma once\n\n')
        f.write('static const unsigned short kMcEdgeMask[256] = {\n')
        for i in range(0, 256, 16):
            f.write('    ' + ', '.join(f'0x{edge_mask(c):03x}' for c in range(i, i + 16)) + ',\n')
        f.write('};\n\nstatic const unsigned char kMcTriCount[256] = {\n')
        for i in range(0, 256, 32):
            f.write('    ' + ', '.join(str(len(rows[c]) // 3) for c in range(i, i + 32)) + ',\n')
        f.write('};\n\n// up to 5 triangles = 15 edge ids per case, padded with 255\nstatic const unsigned char kMcTriTable[256][16] = {\n')
        for c in range(256):
            r = rows[c] + [255] * (16 - len(rows[c]))
            f.write('    {' + ', '.join(f'{v:3d}' for v in r) + '},\n')
        f.write('};\n')


if __name__ == '__main__':
    rows = tri_rows()
    check(rows)
    if '--check' not in sys.argv:
        out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'animatable_nerf_b200', 'csrc', 'mc_tables.h')
        emit(rows, out)
        print('wrote', out)
