"""Generate the 256-case marching-cubes triangle table used by csrc/volume.cu and oracle/volume_oracle.c.

The table is DERIVED here, not transcribed: for every sign configuration of the 8 cube corners

  * on each of the 6 cube faces the crossing points of the face's 4 edges are joined by directed segments
    (walking the face counter-clockwise seen from outside the cube, a segment runs from the crossing where the walk
    leaves the inside set to the next crossing where it re-enters it); on an ambiguous face (corners alternate
    in/out/in/out) the two segments each cut off ONE INSIDE corner -- the rule depends only on the face's own corner
    signs, so the two cubes sharing a face draw the same segments and the surface is watertight;
  * every crossing point is the head of exactly one segment and the tail of exactly one (a cube edge belongs to two
    faces which walk it in opposite directions), so the segments close into oriented loops;
  * each loop is rotated to start at its lowest edge id and fan-triangulated.

Conventions (shared with the kernels):
  corner c  = offsets (c & 1, (c >> 1) & 1, (c >> 2) & 1) on grid axes (0, 1, 2)   [toolpath_utils.py:990-993 order]
  edge  e   = 4 * axis + k: the edge parallel to `axis` whose low corner has the two OTHER offsets (lower axis first)
              equal to (k & 1, k >> 1)
  case bit c is set iff value[c] < level ("inside")
  triangles are wound so that the right-hand normal points from inside to outside (towards increasing values).

usage: python scripts/gen_mc_table.py  -> writes inf-3dp_b200/csrc/mc_table.inc
"""
import itertools
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "inf-3dp_b200", "csrc", "mc_table.inc")


def corner_offset(c):
    return np.array([c & 1, (c >> 1) & 1, (c >> 2) & 1])


def corner_id(o):
    return int(o[0]) | (int(o[1]) << 1) | (int(o[2]) << 2)


def edge_corners(e):
    axis, k = divmod(e, 4)
    others = [a for a in range(3) if a != axis]
    lo = np.zeros(3, dtype=int)
    lo[others[0]] = k & 1
    lo[others[1]] = k >> 1
    hi = lo.copy()
    hi[axis] = 1
    return corner_id(lo), corner_id(hi)


EDGE_OF = {}
for _e in range(12):
    _a, _b = edge_corners(_e)
    EDGE_OF[(_a, _b)] = _e
    EDGE_OF[(_b, _a)] = _e


def face_cycles():
    """The 6 faces as corner 4-cycles, counter-clockwise seen from outside the cube."""
    faces = []
    for axis in range(3):
        for side in (0, 1):
            u, v = [a for a in range(3) if a != axis]
            # (u, v, axis) right-handed?  normal = +axis when side == 1
            eu, ev, en = np.eye(3, dtype=int)[u], np.eye(3, dtype=int)[v], np.eye(3, dtype=int)[axis]
            right_handed = np.dot(np.cross(eu, ev), en) > 0
            base = en * side
            cyc = [base, base + eu, base + eu + ev, base + ev]
            # counter-clockwise about the outward normal (+axis for side 1, -axis for side 0)
            ccw_about_plus = right_handed
            if (side == 1) != ccw_about_plus:
                cyc = cyc[::-1]
            faces.append([corner_id(c) for c in cyc])
    return faces


FACES = face_cycles()


def case_loops(case):
    inside = [(case >> c) & 1 for c in range(8)]
    nxt = {}
    for cyc in FACES:
        outs, ins = [], []   # crossings where the walk leaves / enters the inside set, in walk order
        for i in range(4):
            a, b = cyc[i], cyc[(i + 1) % 4]
            if inside[a] and not inside[b]:
                outs.append((i, EDGE_OF[(a, b)]))
            elif not inside[a] and inside[b]:
                ins.append((i, EDGE_OF[(a, b)]))
        if len(outs) == 1:
            nxt_edge = {outs[0][1]: ins[0][1]}
        elif len(outs) == 2:
            # ambiguous face: cut off each inside corner on its own.  The inside corner cyc[i] is entered through the
            # crossing on walk edge i-1 (an "in" crossing) and left through walk edge i (an "out" crossing): the
            # segment that isolates it joins those two.  Directed from the "out" crossing to the "in" crossing.
            nxt_edge = {}
            for i_out, e_out in outs:
                i_in = (i_out - 1) % 4
                e_in = [e for (i, e) in ins if i == i_in]
                assert len(e_in) == 1
                nxt_edge[e_out] = e_in[0]
        else:
            assert len(outs) == 0
            nxt_edge = {}
        for k, v in nxt_edge.items():
            assert k not in nxt
            nxt[k] = v
    # every crossing edge is the tail of one segment and the head of one
    crossing = [e for e in range(12) if inside[edge_corners(e)[0]] != inside[edge_corners(e)[1]]]
    assert sorted(nxt.keys()) == crossing and sorted(nxt.values()) == crossing, (case, nxt, crossing)
    loops, seen = [], set()
    for e in crossing:
        if e in seen:
            continue
        loop = [e]
        seen.add(e)
        while nxt[loop[-1]] != e:
            loop.append(nxt[loop[-1]])
            seen.add(loop[-1])
        loops.append(loop)
    return loops


def edge_mid(e):
    a, b = edge_corners(e)
    return (corner_offset(a) + corner_offset(b)) / 2.0


def build():
    # orientation check on the one-corner case: the loop around inside corner 0 must give a normal pointing away from it
    loop = case_loops(1)[0]
    p = [edge_mid(e) for e in loop]
    flip = np.dot(np.cross(p[1] - p[0], p[2] - p[0]), np.array([1.0, 1.0, 1.0])) < 0
    table, max_t = [], 0
    for case in range(256):
        tris = []
        for loop in case_loops(case):
            if flip:
                loop = [loop[0]] + loop[:0:-1]
            for i in range(1, len(loop) - 1):
                tris.append((loop[0], loop[i], loop[i + 1]))
        max_t = max(max_t, len(tris))
        table.append(tris)
    return table, max_t


def verify(table):
    """Per-case sanity: complement symmetry of the vertex sets, orientation of every triangle."""
    for case, tris in enumerate(table):
        inside = [(case >> c) & 1 for c in range(8)]
        for t in tris:
            p = [edge_mid(e) for e in t]
            n = np.cross(p[1] - p[0], p[2] - p[0])
            g = np.zeros(3)
            for e in t:
                a, b = edge_corners(e)
                d = corner_offset(b) - corner_offset(a)
                g += d if inside[a] else -d
            assert np.dot(n, g) >= -1e-12, (case, t)
        used = sorted({e for t in tris for e in t})
        crossing = [e for e in range(12) if inside[edge_corners(e)[0]] != inside[edge_corners(e)[1]]]
        assert used == crossing, case


def main():
    table, max_t = build()
    verify(table)
    with open(OUT, "w") as f:
        f.write("// GENERATED by scripts/gen_mc_table.py -- do not edit.  Conventions in that script's docstring.\n")
        f.write(f"#define INF3DP_MC_MAX_TRIS {max_t}\n")
        f.write("#ifndef INF3DP_MC_ARRAY\n#define INF3DP_MC_ARRAY(type, name) static const type name\n#endif\n")
        f.write("INF3DP_MC_ARRAY(unsigned char, INF3DP_MC_NTRI)[256] = {\n")
        for r in range(0, 256, 32):
            f.write("    " + ", ".join(str(len(table[c])) for c in range(r, r + 32)) + ",\n")
        f.write("};\n")
        f.write(f"INF3DP_MC_ARRAY(signed char, INF3DP_MC_TRI)[256][{3 * max_t}] = {{\n")
        for c in range(256):
            flat = list(itertools.chain.from_iterable(table[c]))
            flat += [-1] * (3 * max_t - len(flat))
            f.write("    {" + ", ".join(f"{v:2d}" for v in flat) + "},\n")
        f.write("};\n")
    print(f"wrote {OUT}: max {max_t} triangles per cube, {sum(len(t) for t in table)} triangles over all cases")


if __name__ == "__main__":
    main()
