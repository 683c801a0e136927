"""Marching-cubes case table, GENERATED (not transcribed): for each of the 256 sign configurations of a cube's corners the iso-surface
polygons are found by tracing the crossing points around the cube's faces and fan-triangulated.

Conventions (shared by csrc/mcubes.cu and oracle/mcubes_oracle.py):
  corner c in 0..7 sits at offset (c & 1, (c >> 1) & 1, (c >> 2) & 1) = (x, y, z);  bit c of the case index is set when value(c) > iso;
  edge e = 4 * axis + (b1 + 2 * b2): the edge parallel to `axis` whose other two coordinates (in increasing axis order) are b1, b2;
  its end points are the corners with the axis bit 0 / 1.
Face rule for the ambiguous faces (two diagonal corners inside): the two INSIDE corners are cut off separately.  The rule depends only
on the four corner signs of the face, so the two cubes that share a face agree and the surface is watertight.
Orientation: triangle normals (right-hand rule) point from the inside (value > iso) to the outside.

TRI_TABLE [256, 16] int8: up to 5 triangles as edge triples, -1 terminated.  EDGE_CORNERS [12, 2].  NUM_TRIS [256].
"""
import numpy as np


def corner_offset(c):
    return (c & 1, (c >> 1) & 1, (c >> 2) & 1)


def _edge_id(axis, b1, b2):
    return 4 * axis + (b1 + 2 * b2)


def _edge_corners():
    ec = np.zeros((12, 2), np.int64)
    for axis in range(3):
        others = [a for a in range(3) if a != axis]
        for b2 in range(2):
            for b1 in range(2):
                off = [0, 0, 0]
                off[others[0]], off[others[1]] = b1, b2
                c0 = off[0] | (off[1] << 1) | (off[2] << 2)
                c1 = c0 | (1 << axis)
                ec[_edge_id(axis, b1, b2)] = (c0, c1)
    return ec


EDGE_CORNERS = _edge_corners()
_EDGE_OF = {(int(a), int(b)): e for e, (a, b) in enumerate(EDGE_CORNERS)}
_EDGE_OF.update({(b, a): e for (a, b), e in list(_EDGE_OF.items())})


def _faces():
    """six faces as cyclic corner quadruples"""
    faces = []
    for axis in range(3):
        u, v = [a for a in range(3) if a != axis]
        for side in range(2):
            quad = []
            for du, dv in ((0, 0), (1, 0), (1, 1), (0, 1)):
                off = [0, 0, 0]
                off[axis], off[u], off[v] = side, du, dv
                quad.append(off[0] | (off[1] << 1) | (off[2] << 2))
            faces.append(quad)
    return faces


_FACES = _faces()


def _case(mask):
    inside = [(mask >> c) & 1 for c in range(8)]
    # segments between crossing edges, face by face
    adj = {}

    def link(e0, e1):
        adj.setdefault(e0, []).append(e1)
        adj.setdefault(e1, []).append(e0)

    for quad in _FACES:
        cross = []                                  # (position in the cycle, edge) of the face's crossed edges
        for k in range(4):
            a, b = quad[k], quad[(k + 1) % 4]
            if inside[a] != inside[b]:
                cross.append((k, _EDGE_OF[(a, b)]))
        if len(cross) == 2:
            link(cross[0][1], cross[1][1])
        elif len(cross) == 4:
            # corners alternate inside / outside around the face: cut off each INSIDE corner (quad[k] is between edge k-1 and edge k)
            for k in range(4):
                if inside[quad[k]]:
                    link(_EDGE_OF[(quad[(k - 1) % 4], quad[k])], _EDGE_OF[(quad[k], quad[(k + 1) % 4])])
    # every crossed edge lies on two faces: degree 2, disjoint cycles
    tris = []
    seen = set()
    mid = {e: (np.array(corner_offset(int(EDGE_CORNERS[e, 0])), float) + np.array(corner_offset(int(EDGE_CORNERS[e, 1])), float)) / 2 for e in adj}
    for start in sorted(adj):
        if start in seen:
            continue
        assert len(adj[start]) == 2
        loop = [start]
        seen.add(start)
        prev, cur = start, adj[start][0]
        while cur != start:
            loop.append(cur)
            seen.add(cur)
            nxt = [n for n in adj[cur] if n != prev]
            if len(nxt) == 2:                     # both links go to the same edge (a 2-cycle cannot occur on a cube)
                nxt = [adj[cur][1]]
            prev, cur = cur, (nxt[0] if nxt else start)
        assert len(loop) >= 3
        # orientation: the polygon's normal must point from the inside end of its edges to the outside end
        n = np.zeros(3)
        ctr = np.mean([mid[e] for e in loop], axis=0)
        for k in range(len(loop)):
            n += np.cross(mid[loop[k]] - ctr, mid[loop[(k + 1) % len(loop)]] - ctr)
        s = 0.0
        for e in loop:
            c0, c1 = (int(x) for x in EDGE_CORNERS[e])
            d = np.array(corner_offset(c1), float) - np.array(corner_offset(c0), float)
            s += np.dot(n, d if inside[c0] else -d)
        if s < 0:
            loop = loop[::-1]
        for k in range(1, len(loop) - 1):
            tris.append((loop[0], loop[k], loop[k + 1]))
    return tris


def _build():
    table = np.full((256, 16), -1, np.int8)
    ntri = np.zeros(256, np.int32)
    for mask in range(256):
        tris = _case(mask)
        assert len(tris) <= 5, (mask, len(tris))
        ntri[mask] = len(tris)
        for t, tri in enumerate(tris):
            table[mask, 3 * t:3 * t + 3] = tri
    return table, ntri


TRI_TABLE, NUM_TRIS = _build()
