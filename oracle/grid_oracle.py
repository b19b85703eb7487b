"""TEST INFRASTRUCTURE ONLY. CPU restatement of the reference's grid finder, find_grid.cc:1216-1445
(mrgingham::find_grid_from_points), for checking the library's host-side grid finder.

PARITY: pinned by tests/test_grid_vs_ref.py against the reference's own find_grid.cc, compiled unmodified
over a stand-in for Boost.Polygon's voronoi_diagram (oracle/shim/boost/polygon/voronoi.hpp; Boost itself is
not in this image and the reference ships no golden vectors for the grid finder), so what stays MODELLED
is only the neighbour graph Boost would hand over -- the conventions listed below, which the stand-in
shares. What is restated exactly is everything the reference itself computes on top of that graph (the neighbour walk, find_grid.cc:86-140; the sequence search,
:207-343; the outer-edge cycles, :776-1013; the orientation logic, :1015-1187; the row fill, :1385-1439).
What stands in for Boost here:
  * Two sites are neighbours iff their Voronoi cells share an edge of non-zero length. This module decides
    that from the definition (an empty circle through both points whose centre can move along their
    bisector), in exact integer arithmetic, independently of the triangulation the library builds. Boost
    removes zero-length edges between cocircular sites as well (voronoi_diagram.hpp, "remove degenerate edges").
  * Cells are visited in the order Boost creates them: sites sorted by (x, y).
  * Around a cell the edges run counter-clockwise in (x, y) (clockwise as seen in an image with y down,
    find_grid.cc:40-41). The edge Boost starts at (cell->incident_edge()) depends on its sweep-line
    internals; here the walk starts at the first neighbour counter-clockwise from the +x direction. The
    start only matters where several neighbours satisfy the reference's "first match wins" rule
    (find_grid.cc:216-221), which the reference declares out of contract ("assuming clean data").
"""
import math
from fractions import Fraction

import numpy as np

FIND_GRID_SCALE = 1000                      # mrgingham-internal.h:3
FIND_GRID_SCALE_APPROX_POWER2 = 1024        # mrgingham-internal.h:6
THRESHOLD_SPACING_COS = 0.984               # find_grid.cc:202-205
THRESHOLD_SPACING_LENGTH_RATIO_MIN = 0.7
THRESHOLD_SPACING_LENGTH_RATIO_MAX = 1.4
THRESHOLD_SPACING_LENGTH_RATIO_DEVIATION = 0.35


def _cdiv(a, b):
    """C integer division (truncates toward zero)"""
    q = abs(a) // abs(b)
    return q if (a >= 0) == (b >= 0) else -q


def _fdiv(a, b):
    """IEEE double division, including x/0 and 0/0"""
    if b == 0.0:
        if a == 0.0 or a != a:
            return float("nan")
        return math.copysign(float("inf"), a) * math.copysign(1.0, b)
    return a / b


def _angle_key_less(a, b):
    """a, b: integer vectors; True if a comes before b counter-clockwise starting at the +x direction"""
    ha = 0 if (a[1] > 0 or (a[1] == 0 and a[0] > 0)) else 1
    hb = 0 if (b[1] > 0 or (b[1] == 0 and b[0] > 0)) else 1
    if ha != hb:
        return ha < hb
    return a[0] * b[1] - a[1] * b[0] > 0


def voronoi_neighbours(points):
    """points: list of (x,y) Python ints, distinct. Returns for each point the list of points whose Voronoi
    cells share an edge of non-zero length with its cell, by the definition: a and b are such neighbours
    iff the centres t on their bisector (centre = (a+b)/2 + t*perp(b-a)) whose circle through a and b has
    every other point strictly outside form an interval of non-zero length."""
    n = len(points)
    P = np.array(points, dtype=np.int64).reshape(n, 2)
    nb = [[] for _ in range(n)]
    for ia in range(n):
        ax, ay = points[ia]
        for ib in range(ia + 1, n):
            bx, by = points[ib]
            dx, dy = -(by - ay), bx - ax                       # direction of the bisector
            # point c is strictly outside the circle centred at m + t d  <=>  t * alpha_c > beta_c, with
            # alpha = 2 d.(a-c), beta = |a|^2-|c|^2 - (a+b).(a-c)      (all integers, < 2^56 for 25-bit input)
            acx, acy = ax - P[:, 0], ay - P[:, 1]
            alpha = 2 * (dx * acx + dy * acy)
            beta = (ax * ax + ay * ay) - (P[:, 0] ** 2 + P[:, 1] ** 2) - ((ax + bx) * acx + (ay + by) * acy)
            keep = np.ones(n, bool); keep[ia] = False; keep[ib] = False
            al, be = alpha[keep], beta[keep]
            zero = al == 0
            if np.any(zero & (be >= 0)):                        # a point on the segment's line between a and b
                continue
            pos, neg = al > 0, al < 0
            # float screen, exact decision near a tie
            lo = (be[pos] / al[pos]).max() if pos.any() else -np.inf
            hi = (be[neg] / al[neg]).min() if neg.any() else np.inf
            if lo < hi and not (np.isfinite(lo) and np.isfinite(hi) and (hi - lo) <= 1e-6 * max(1.0, abs(lo), abs(hi))):
                ok = True
            elif lo > hi and (lo - hi) > 1e-6 * max(1.0, abs(lo), abs(hi)):
                ok = False
            else:
                lo_e = max(Fraction(int(b_), int(a_)) for a_, b_ in zip(al[pos], be[pos]))
                hi_e = min(Fraction(int(b_), int(a_)) for a_, b_ in zip(al[neg], be[neg]))
                ok = lo_e < hi_e
            if ok:
                nb[ia].append(ib); nb[ib].append(ia)
    return nb


class _Graph:
    """the part of Boost's voronoi_diagram the reference uses: cells in site order, and around each cell
    the neighbouring cells in counter-clockwise order"""

    def __init__(self, points):
        # Boost sorts the sites and drops duplicates; a cell's source_index is an index into `points`
        self.pts = [(int(p[0]), int(p[1])) for p in points]
        first = {}
        for i, p in enumerate(self.pts):
            first.setdefault(p, i)
        self.sites = sorted(first.values(), key=lambda i: self.pts[i])
        upts = [self.pts[i] for i in self.sites]
        nb = voronoi_neighbours(upts)
        self.ring = {}
        for k, i in enumerate(self.sites):
            vecs = [(self.sites[j], (upts[j][0] - upts[k][0], upts[j][1] - upts[k][1])) for j in nb[k]]
            # insertion sort with the exact comparator
            out = []
            for item in vecs:
                pos = 0
                while pos < len(out) and _angle_key_less(out[pos][1], item[1]):
                    pos += 1
                out.insert(pos, item)
            self.ring[i] = [j for j, _ in out]

    def next(self, a, b):
        r = self.ring[a]
        return r[(r.index(b) + 1) % len(r)]

    def prev(self, a, b):
        r = self.ring[a]
        return r[(r.index(b) - 1) % len(r)]

    def adjacent(self, a):
        """find_grid.cc:86-140 (FOR_ALL_ADJACENT_CELLS): every edge-neighbour, each followed by the cell
        'in between' this neighbour and the next one where that cell exists and lies angularly between them.
        Yields (cell, delta)."""
        P = self.pts
        for b in list(self.ring[a]):
            yield b, (P[b][0] - P[a][0], P[b][1] - P[a][1])
            c = self.next(a, b)
            v0 = (P[b][0] - P[a][0], P[b][1] - P[a][1])
            v1 = (P[c][0] - P[a][0], P[c][1] - P[a][1])
            if v1[0] * v0[1] > v0[0] * v1[1]:
                continue
            if self.prev(b, a) != c:
                continue
            d = self.prev(b, c)
            vm = (P[d][0] - P[a][0], P[d][1] - P[a][1])
            if v1[0] * vm[1] > vm[0] * v1[1]:
                continue
            if vm[0] * v0[1] > v0[0] * vm[1]:
                continue
            yield d, vm


class _Stats:
    def __init__(self, delta):
        self.delta_last = delta
        self.ratio_sum = 0.0
        self.ratio_n = 0


def _next_along_sequence(g, stats, c):
    """find_grid.cc:207-310"""
    dl = stats.delta_last
    dl_len = math.hypot(float(dl[0]), float(dl[1]))
    for c_adj, delta in g.adjacent(c):
        d_len = math.hypot(float(delta[0]), float(delta[1]))
        cos_err = _fdiv(float(dl[0]) * float(delta[0]) + float(dl[1]) * float(delta[1]), dl_len * d_len)
        if cos_err < THRESHOLD_SPACING_COS:
            continue
        ratio = _fdiv(d_len, dl_len)
        if ratio < THRESHOLD_SPACING_LENGTH_RATIO_MIN or ratio > THRESHOLD_SPACING_LENGTH_RATIO_MAX:
            continue
        if stats.ratio_n > 2:
            dev = ratio - stats.ratio_sum / float(stats.ratio_n)
            if dev < -THRESHOLD_SPACING_LENGTH_RATIO_DEVIATION or dev > THRESHOLD_SPACING_LENGTH_RATIO_DEVIATION:
                continue
        stats.ratio_sum += ratio
        stats.ratio_n += 1
        stats.delta_last = delta
        return c_adj
    return None


def _walk(g, delta, c, n_remaining):
    """the cells FOR_MATCHING_ADJACENT_CELLS visits (find_grid.cc:187-197); None where the walk breaks"""
    stats = _Stats(delta)
    out = []
    for _ in range(n_remaining):
        c = _next_along_sequence(g, stats, c)
        out.append((c, stats.delta_last))
        if c is None:
            break
    return out


def _search_along_sequence(g, delta, c, n_remaining):
    """find_grid.cc:312-343; returns (clast, delta_mean) or None"""
    mx, my = float(delta[0]), float(delta[1])
    clast = None
    for c_adj, dlast in _walk(g, delta, c, n_remaining):
        if c_adj is None:
            return None
        mx += float(dlast[0]); my += float(dlast[1])
        clast = c_adj
    if clast is None:
        return None
    return clast, (mx / float(n_remaining + 1), my / float(n_remaining + 1))


def _sequence_points(g, cs, gridn):
    """find_grid.cc:579-613"""
    c0, c1 = cs["c0"], cs["c1"]
    delta = (g.pts[c1][0] - g.pts[c0][0], g.pts[c1][1] - g.pts[c0][1])
    return [c0, c1] + [c for c, _ in _walk(g, delta, c1, gridn - 2)]


def _is_crossing(P, a0, a1, b0, b1):
    """find_grid.cc:776-822, in float32 as the reference computes it"""
    f = np.float32
    l0 = (f(P[a1][0] - P[a0][0]), f(P[a1][1] - P[a0][1]))
    p0 = (f(P[b0][0] - P[a0][0]), f(P[b0][1] - P[a0][1]))
    p1 = (f(P[b1][0] - P[a0][0]), f(P[b1][1] - P[a0][1]))
    with np.errstate(all="ignore"):
        d2 = f(f(l0[0] * l0[0]) + f(l0[1] * l0[1]))
        r0 = (f(f(p0[0] * l0[0]) + f(p0[1] * l0[1])), f(f(f(-p0[0]) * l0[1]) + f(p0[1] * l0[0])))
        r1 = (f(f(p1[0] * l0[0]) + f(p1[1] * l0[1])), f(f(f(-p1[0]) * l0[1]) + f(p1[1] * l0[0])))
        if f(r0[1] * r1[1]) > 0:
            return False
        if (r0[0] < 0 and r1[0] < 0) or (r0[0] > d2 and r1[0] > d2):
            return False
        k = f(r0[1] / f(r0[1] - r1[1]))
        x = f(r0[0] + f(k * f(r1[0] - r0[0])))
        return bool(x >= f(0) and x <= d2)


def find_grid_from_points(points, gridn):
    """points: int array [n,2], scaled by 1000 (PointInt). Returns float64 [gridn*gridn, 2] in pixels, in the
    reference's order (rows from the top edge, find_grid.cc:1364-1439), or None."""
    points = np.asarray(points).reshape(-1, 2)
    if len(points) == 0:
        return None
    g = _Graph(points)
    P = g.pts

    # every sequence of gridn cells, from every cell in every direction (find_grid.cc:505-566)
    cands = []
    for c in g.sites:
        for c_adj, delta in g.adjacent(c):
            r = _search_along_sequence(g, delta, c_adj, gridn - 2)
            if r is not None:
                cands.append({"c0": c, "c1": c_adj, "clast": r[0], "delta_mean": r[1]})

    # outer edges: sequences that start at a cell starting at least two sequences (:1244-1275)
    count = {}
    for cs in cands:
        count[cs["c0"]] = count.get(cs["c0"], 0) + 1
    outer = [i for i, cs in enumerate(cands) if count[cs["c0"]] >= 2]
    if len(outer) < 8:
        return None
    from_point = {}
    for i, ics in enumerate(outer):
        from_point.setdefault(cands[ics]["c0"], []).append(i)

    def first(i): return cands[outer[i]]["c0"]
    def last(i): return cands[outer[i]]["clast"]

    def next_outer_edge(edges, edge_count, point_initial):
        """find_grid.cc:826-960: extend edges[:edge_count] to a unique 4-cycle back to point_initial"""
        found = None
        i_edge = edges[edge_count - 1]
        nxt = from_point.get(last(i_edge))
        if nxt is None:
            return False
        for e in nxt:
            if last(e) == first(i_edge):
                continue
            if edge_count != 3:
                if last(e) == point_initial:
                    continue
                if edge_count == 2 and _is_crossing(P, first(edges[0]), last(edges[0]), first(e), last(e)):
                    continue
                edges[edge_count] = e
                if not next_outer_edge(edges, edge_count + 1, point_initial):
                    continue
                if found is not None:
                    return False
                found = list(edges)
            else:
                if last(e) != point_initial:
                    continue
                if _is_crossing(P, first(edges[1]), last(edges[1]), first(e), last(e)):
                    return False
                edges[3] = e
                return True
        if found is None:
            return False
        edges[:] = found
        return True

    cycles, used = [], set()
    for i in range(len(outer)):
        if i in used:
            continue
        edges = [i, 0, 0, 0]
        if not next_outer_edge(edges, 1, first(i)):
            continue
        cycles.append(list(edges))
        used.update(edges)
    if len(cycles) < 2:
        return None

    def equal_and_opposite(c0, c1):
        """find_grid.cc:962-1013"""
        i0, p0 = 0, first(c0[0])
        i1 = -1
        for k in range(4):
            if last(c1[k]) == p0:
                i1 = k
                break
        if i1 < 0:
            return False
        for _ in range(4):
            if first(c0[i0]) != last(c1[i1]) or last(c0[i0]) != first(c1[i1]):
                return False
            i0 = (i0 + 1) % 4
            i1 = (i1 + 3) % 4
        return True

    pair = None
    for a in range(len(cycles)):
        for b in range(a + 1, len(cycles)):
            if equal_and_opposite(cycles[a], cycles[b]):
                if pair is not None:
                    return None
                pair = (a, b)
    if pair is None:
        return None

    # which of the two is clockwise, and which edge of each is the top (:1015-1187)
    S = FIND_GRID_SCALE_APPROX_POWER2
    cyc = [cycles[pair[0]], cycles[pair[1]]]
    v = [(_cdiv(P[last(e)][0] - P[first(e)][0], S), _cdiv(P[last(e)][1] - P[first(e)][1], S)) for e in cyc[0]]
    sign = [v[(i + 1) % 4][0] * v[i][1] < v[i][0] * v[(i + 1) % 4][1] for i in range(4)]
    if all(sign):
        iclockwise = 0
    elif not any(sign):
        iclockwise = 1
    else:
        return None
    iedge_top = [0, 0]
    for ic in range(2):
        INT_MAX = 2 ** 31 - 1
        ymin, iedge, pmin, pmax = [INT_MAX, INT_MAX], [-1, -1], [0, 0], [0, 0]
        for i in range(4):
            p0, p1 = first(cyc[ic][i]), last(cyc[ic][i])
            if P[p0][1] < P[p1][1]:
                y, lo, hi = P[p0][1], p0, p1
            else:
                y, lo, hi = P[p1][1], p1, p0
            if y < ymin[0]:
                ymin[1], iedge[1], pmin[1], pmax[1] = ymin[0], iedge[0], pmin[0], pmax[0]
                ymin[0], iedge[0], pmin[0], pmax[0] = y, i, lo, hi
            elif y < ymin[1]:
                ymin[1], iedge[1], pmin[1], pmax[1] = y, i, lo, hi
        v0y = _cdiv(P[pmax[0]][1] - P[pmin[0]][1], S); v0x = abs(_cdiv(P[pmax[0]][0] - P[pmin[0]][0], S))
        v1y = _cdiv(P[pmax[1]][1] - P[pmin[1]][1], S); v1x = abs(_cdiv(P[pmax[1]][0] - P[pmin[1]][0], S))
        cross = (v0x * v1y - v0y * v1x) ** 2
        denom = (v0x * v0x + v0y * v0y) * (v1x * v1x + v1y * v1y)
        if abs(cross) * 8 < denom * 1:
            return None
        iedge_top[ic] = iedge[0] if abs(v0y * v1x) < abs(v1y * v0x) else iedge[1]

    # rows: from the i-th cell of the left edge to the i-th cell of the right edge (:1364-1439)
    seq_from = {}
    for i, cs in enumerate(cands):
        seq_from.setdefault(cs["c0"], []).append(i)

    def seq_from_to(a, b):
        for i in seq_from.get(a, []):
            if cands[i]["clast"] == b:
                return i
        return -1

    rows = [outer[cyc[iclockwise][iedge_top[iclockwise]]]]
    vleft = outer[cyc[1 - iclockwise][(iedge_top[1 - iclockwise] + 1) % 4]]
    vright = outer[cyc[iclockwise][(iedge_top[iclockwise] + 1) % 4]]
    lpts = _sequence_points(g, cands[vleft], gridn)
    rpts = _sequence_points(g, cands[vright], gridn)
    for i in range(1, gridn):
        s = seq_from_to(lpts[i], rpts[i])
        if s < 0 or seq_from_to(rpts[i], lpts[i]) < 0:
            return None
        rows.append(s)
    out = []
    for s in rows:
        for c in _sequence_points(g, cands[s], gridn):
            out.append((float(P[c][0]) / float(FIND_GRID_SCALE), float(P[c][1]) / float(FIND_GRID_SCALE)))
    return np.array(out, dtype=np.float64)
