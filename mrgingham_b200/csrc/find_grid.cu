// Host side of SURVEY.md row F1: the grid finder, mrgingham::find_grid_from_points (find_grid.cc:1216-1445),
// which turns the detector's unordered corner list into the ordered gridn x gridn board. A few hundred
// points per frame: this is host code (one call per frame, microseconds), not a kernel.
//
// The reference gets its neighbour graph from Boost.Polygon's voronoi_diagram (find_grid.cc:7,1226). This
// file builds what the reference reads from that diagram from scratch:
//   * an exact Delaunay triangulation of the integer sites (insertion outside the current hull in order of distance
//     from the centroid, Lawson flips, predicates exact: doubles behind a static error bound, 128-bit integers where
//     that cannot decide), whose edges, minus those between cocircular sites
//     (Voronoi edges of zero length, which Boost removes too), are the Voronoi edges;
//   * per cell, the neighbouring cells in counter-clockwise order in (x,y) (find_grid.cc:40-41: clockwise as
//     seen in an image), cells visited in sorted-site order as Boost creates them.
// Parity: Boost is not in this image; tests/test_grid_vs_ref.py compares this file with the reference's own
// find_grid.cc compiled over a stand-in for Boost's voronoi_diagram that fixes the same conventions. The one
// thing that stays modelled is which edge Boost starts a cell's walk at; it matters only when several
// neighbours pass the reference's "first match wins" test (find_grid.cc:216-221).
// Everything on top of the graph follows the reference's arithmetic (doubles, the float32 crossing test,
// the truncating integer divisions).
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <sys/stat.h>
#include <map>
#include <set>
#include <string>
#include <vector>

#include "find_grid.hh"

namespace mrgb200
{
namespace
{
typedef long long i64;
typedef __int128  i128;

constexpr int    kScale = 1000;        // FIND_GRID_SCALE, mrgingham-internal.h:3
constexpr int    kScalePow2 = 1024;    // FIND_GRID_SCALE_APPROX_POWER2, mrgingham-internal.h:6
constexpr double kMinCos = 0.984, kMinRatio = 0.7, kMaxRatio = 1.4, kMaxRatioDeviation = 0.35;   // find_grid.cc:202-205
constexpr i64    kMaxCoord = 1ll << 29;   // keeps the in-circle determinant inside 128 bits

struct P2 { i64 x, y; };

inline i64 orient(const P2& a, const P2& b, const P2& c)      // > 0: a,b,c counter-clockwise
{
    return (b.x - a.x) * (c.y - a.y) - (b.y - a.y) * (c.x - a.x);
}
// sign of the in-circle determinant: > 0 iff d is strictly inside the circle through the ccw triangle a,b,c.
// The differences are exact in double (|coordinate| <= 2^29); the determinant is evaluated in double first and its sign
// taken when it clears Shewchuk's static error bound for this expression ((10 + 96 eps) eps times the sum of the
// absolute terms); only the near-cocircular cases -- which a regular grid has plenty of -- pay for 128-bit integers.
inline int incircle(const P2& a, const P2& b, const P2& c, const P2& d)
{
    const i64 ax = a.x - d.x, ay = a.y - d.y, bx = b.x - d.x, by = b.y - d.y, cx = c.x - d.x, cy = c.y - d.y;
    {
        const double fax = (double)ax, fay = (double)ay, fbx = (double)bx, fby = (double)by, fcx = (double)cx, fcy = (double)cy;
        const double bxcy = fbx * fcy, cxby = fcx * fby, cxay = fcx * fay, axcy = fax * fcy, axby = fax * fby, bxay = fbx * fay;
        const double al = fax * fax + fay * fay, bl = fbx * fbx + fby * fby, cl = fcx * fcx + fcy * fcy;
        const double det = al * (bxcy - cxby) + bl * (cxay - axcy) + cl * (axby - bxay);
        const double perm = (std::fabs(bxcy) + std::fabs(cxby)) * al + (std::fabs(cxay) + std::fabs(axcy)) * bl + (std::fabs(axby) + std::fabs(bxay)) * cl;
        const double bound = 1.2e-15 * perm;          // (10 + 96 * 2^-53) * 2^-53 = 1.11e-15
        if (det > bound) return 1;
        if (det < -bound) return -1;
    }
    const i128 a2 = (i128)ax * ax + (i128)ay * ay, b2 = (i128)bx * bx + (i128)by * by, c2 = (i128)cx * cx + (i128)cy * cy;
    const i128 det = a2 * ((i128)bx * cy - (i128)by * cx) - b2 * ((i128)ax * cy - (i128)ay * cx) + c2 * ((i128)ax * by - (i128)ay * bx);
    return det > 0 ? 1 : det < 0 ? -1 : 0;
}
// true if a comes before b going counter-clockwise from the +x direction
inline bool angle_less(const P2& a, const P2& b)
{
    const int ha = (a.y > 0 || (a.y == 0 && a.x > 0)) ? 0 : 1, hb = (b.y > 0 || (b.y == 0 && b.x > 0)) ? 0 : 1;
    if (ha != hb) return ha < hb;
    return a.x * b.y - a.y * b.x > 0;
}

// ---- the neighbour graph ----
struct Graph
{
    std::vector<P2>  pts;        // by source index
    std::vector<int> sites;      // distinct sites, sorted by (x,y): the order cells are visited in
    // counter-clockwise ring of edge-neighbours of every site (by source index)
    std::vector<int> ring_off, ring;
    // the walk of find_grid.cc:86-140 precomputed: each edge-neighbour followed by the in-between cell, if any
    std::vector<int> adj_off, adj;
    std::vector<double> adj_dx, adj_dy, adj_len;   // the step to each of them and its hypot()
    // Per step (entry of adj: a -> b), the steps out of b that may follow it in a sequence: those that pass the
    // direction and length-ratio tests of find_grid.cc:247-284, which depend on the two steps only. In adj order
    // ("first match wins", find_grid.cc:216-221), with the length ratio each one contributes.
    std::vector<int> cont_off, cont;
    std::vector<double> cont_ratio;

    int ring_pos(int a, int b) const
    {
        for (int k = ring_off[a]; k < ring_off[a + 1]; k++) if (ring[k] == b) return k;
        return -1;
    }
    int prev(int a, int b) const { const int k = ring_pos(a, b); return ring[k == ring_off[a] ? ring_off[a + 1] - 1 : k - 1]; }
};

struct Tri { int v[3]; int n[3]; };

class Triangulation
{
public:
    Triangulation(const std::vector<P2>& pts, const std::vector<int>& sorted_sites) : P(pts), sorted(sorted_sites) {}

    // directed neighbour pairs (a,b), each neighbour relation in both directions; false if no triangle exists
    // (all sites on one line)
    bool run(std::vector<std::pair<int,int>>* nb)
    {
        const int n = (int)sorted.size();
        // Insertion order: by distance from (a lattice point near) the centroid, all in exact integers. Every site is then
        // outside the hull of the sites before it (they lie in the closed disc it is on the rim of, or beyond), and the
        // triangles it adds are close to Delaunay already: a fraction of the flips that sites sorted by x need on a
        // regular grid, where each new site first sees a whole column of thin triangles.
        ord = sorted;
        {
            i64 sx = 0, sy = 0;
            for (int k = 0; k < n; k++) { sx += P[ord[k]].x; sy += P[ord[k]].y; }
            const i64 cx = sx / n, cy = sy / n;
            std::vector<std::pair<i64, int>> key(n);
            for (int k = 0; k < n; k++)
            {
                const i64 dx = P[ord[k]].x - cx, dy = P[ord[k]].y - cy;
                key[k] = std::make_pair(dx * dx + dy * dy, k);         // (ties: in sorted-site order)
            }
            std::sort(key.begin(), key.end());
            for (int k = 0; k < n; k++) ord[k] = sorted[key[k].second];
        }
        int m = 2;
        while (m < n && orient(P[ord[0]], P[ord[1]], P[ord[m]]) == 0) m++;
        if (m >= n) return false;
        // the sites before the first one off their common line, in order along the line (seed() fans them out from it)
        std::sort(ord.begin(), ord.begin() + m, [&](int a, int b) { return P[a].x != P[b].x ? P[a].x < P[b].x : P[a].y < P[b].y; });
        hnext.assign(P.size(), -1); hprev.assign(P.size(), -1); htri.assign(P.size(), -1);
        T.reserve(2 * n);
        seed(m);
        for (int i = m + 1; i < n; i++) if (!insert(ord[i])) return false;

        nb->clear();
        nb->reserve(3 * T.size() + 16);
        for (size_t t = 0; t < T.size(); t++)
            for (int i = 0; i < 3; i++)
            {
                // directed edge a -> b, opposite vertex c
                const int a = T[t].v[(i + 1) % 3], b = T[t].v[(i + 2) % 3], c = T[t].v[i], u = T[t].n[i];
                if (u < 0) { nb->push_back(std::make_pair(a, b)); nb->push_back(std::make_pair(b, a)); continue; }
                if (u < (int)t) continue;          // an inner edge is looked at once, from the lower-numbered triangle
                int j = 0; while (T[u].n[j] != (int)t) j++;
                // the Voronoi edge between a and b has zero length when the two triangles share a circumcircle
                if (incircle(P[c], P[a], P[b], P[T[u].v[j]]) == 0) continue;
                nb->push_back(std::make_pair(a, b)); nb->push_back(std::make_pair(b, a));
            }
        return true;
    }

private:
    const std::vector<P2>& P;
    const std::vector<int>& sorted;          // distinct sites, sorted by (x,y)
    std::vector<int> ord;                     // insertion order
    std::vector<Tri> T;
    std::vector<int> hnext, hprev, htri;      // convex hull, counter-clockwise; htri[v] owns the edge v -> hnext[v]
    int last = -1;                            // most recently inserted site (always a hull corner)
    std::vector<std::pair<int,int>> stack;

    int add(int a, int b, int c) { Tri t; t.v[0] = a; t.v[1] = b; t.v[2] = c; t.n[0] = t.n[1] = t.n[2] = -1; T.push_back(t); return (int)T.size() - 1; }
    void link_hull(int a, int b, int t) { hnext[a] = b; hprev[b] = a; htri[a] = t; }
    // make t's edge (a -> b) and u's edge (b -> a) neighbours
    void glue(int t, int u, int a, int b)
    {
        for (int i = 0; i < 3; i++)
        {
            if (T[t].v[(i + 1) % 3] == a && T[t].v[(i + 2) % 3] == b) T[t].n[i] = u;
            if (T[u].v[(i + 1) % 3] == b && T[u].v[(i + 2) % 3] == a) T[u].n[i] = t;
        }
    }

    // ord[0..m-1] lie on one line (in sorted order along it); ord[m] is off it: a fan
    void seed(int m)
    {
        const int q = ord[m];
        const bool left = orient(P[ord[0]], P[ord[1]], P[q]) > 0;
        int prev_t = -1;
        for (int i = 0; i + 1 < m; i++)
        {
            const int a = ord[i], b = ord[i + 1];
            const int t = left ? add(a, b, q) : add(b, a, q);
            if (left) link_hull(a, b, t); else link_hull(b, a, t);
            if (prev_t >= 0) { if (left) glue(prev_t, t, a, q); else glue(prev_t, t, q, a); }
            prev_t = t;
        }
        if (left) { link_hull(ord[m - 1], q, prev_t); link_hull(q, ord[0], 0); }
        else      { link_hull(q, ord[m - 1], prev_t); link_hull(ord[0], q, 0); }
        last = q;
    }

    // p is at least as far from the centre of the insertion order as every inserted site, so it is strictly outside
    // their hull: some hull edge has p strictly on its outer side. Found by walking the hull from the last inserted
    // site (a few dozen corners at most on a few hundred sites), then widened to the whole visible chain lo .. hi.
    bool insert(int p)
    {
        int v = last, guard = (int)P.size() + 1;
        while (orient(P[v], P[hnext[v]], P[p]) >= 0) { v = hnext[v]; if (--guard < 0) return false; }
        int lo = v, hi = hnext[v];
        while (orient(P[hi], P[hnext[hi]], P[p]) < 0) hi = hnext[hi];
        while (orient(P[hprev[lo]], P[lo], P[p]) < 0) lo = hprev[lo];
        if (lo == hi) return false;     // cannot happen for a site outside the hull
        int prev_t = -1, first_t = -1;
        for (int a = lo; a != hi; )
        {
            const int b = hnext[a], old = htri[a];
            const int t = add(b, a, p);
            glue(t, old, b, a);
            if (prev_t >= 0) glue(prev_t, t, p, a); else first_t = t;
            prev_t = t;
            stack.push_back(std::make_pair(t, 2));
            a = b;
        }
        link_hull(lo, p, first_t);
        link_hull(p, hi, prev_t);
        last = p;
        legalize();
        return true;
    }

    void legalize()
    {
        while (!stack.empty())
        {
            const int t = stack.back().first, i = stack.back().second;
            stack.pop_back();
            const int u = T[t].n[i];
            if (u < 0) continue;
            int j = 0; while (T[u].n[j] != t) j++;
            const int a = T[t].v[i], b = T[t].v[(i + 1) % 3], c = T[t].v[(i + 2) % 3], d = T[u].v[j];
            if (incircle(P[a], P[b], P[c], P[d]) <= 0) continue;
            // flip the shared edge b-c to a-d: t = (a,b,d), u = (a,d,c)
            const int n_ac = T[t].n[(i + 1) % 3], n_ab = T[t].n[(i + 2) % 3];
            const int n_bd = T[u].n[(j + 1) % 3], n_dc = T[u].n[(j + 2) % 3];
            T[t].v[0] = a; T[t].v[1] = b; T[t].v[2] = d; T[t].n[0] = n_bd; T[t].n[1] = u; T[t].n[2] = n_ab;
            T[u].v[0] = a; T[u].v[1] = d; T[u].v[2] = c; T[u].n[0] = n_dc; T[u].n[1] = n_ac; T[u].n[2] = t;
            if (n_bd >= 0) { for (int k = 0; k < 3; k++) if (T[n_bd].n[k] == u) T[n_bd].n[k] = t; } else htri[b] = t;
            if (n_ac >= 0) { for (int k = 0; k < 3; k++) if (T[n_ac].n[k] == t) T[n_ac].n[k] = u; } else htri[c] = u;
            stack.push_back(std::make_pair(t, 0));
            stack.push_back(std::make_pair(u, 0));
        }
    }
};

bool build_graph(Graph* g, const int* xy, int n)
{
    g->pts.resize(n);
    for (int i = 0; i < n; i++)
    {
        g->pts[i].x = xy[2 * i]; g->pts[i].y = xy[2 * i + 1];
        if (g->pts[i].x < -kMaxCoord || g->pts[i].x > kMaxCoord || g->pts[i].y < -kMaxCoord || g->pts[i].y > kMaxCoord) return false;
    }
    // sorted, distinct sites (Boost sorts the site events and drops repeated ones)
    std::vector<int>& s = g->sites;
    s.resize(n);
    for (int i = 0; i < n; i++) s[i] = i;
    const std::vector<P2>& P = g->pts;
    std::sort(s.begin(), s.end(), [&](int a, int b) { return P[a].x != P[b].x ? P[a].x < P[b].x : P[a].y != P[b].y ? P[a].y < P[b].y : a < b; });
    s.erase(std::unique(s.begin(), s.end(), [&](int a, int b) { return P[a].x == P[b].x && P[a].y == P[b].y; }), s.end());

    std::vector<std::pair<int,int>> nb;
    Triangulation tri(P, s);
    if (!tri.run(&nb))
    {
        // every site on one line: consecutive sites are neighbours
        nb.clear();
        for (size_t k = 0; k + 1 < s.size(); k++) { nb.push_back(std::make_pair(s[k], s[k + 1])); nb.push_back(std::make_pair(s[k + 1], s[k])); }
    }
    g->ring_off.assign(n + 1, 0);
    for (size_t k = 0; k < nb.size(); k++) g->ring_off[nb[k].first + 1]++;
    for (int i = 0; i < n; i++) g->ring_off[i + 1] += g->ring_off[i];
    g->ring.resize(g->ring_off[n]);
    {
        std::vector<int> fill(g->ring_off.begin(), g->ring_off.end() - 1);
        for (size_t k = 0; k < nb.size(); k++) g->ring[fill[nb[k].first]++] = nb[k].second;
    }
    for (int i = 0; i < n; i++)
        std::sort(g->ring.begin() + g->ring_off[i], g->ring.begin() + g->ring_off[i + 1], [&](int a, int b)
                  { const P2 va = { P[a].x - P[i].x, P[a].y - P[i].y }, vb = { P[b].x - P[i].x, P[b].y - P[i].y }; return angle_less(va, vb); });
    // the neighbours the reference looks at from each cell (find_grid.cc:86-140)
    g->adj_off.assign(n + 1, 0);
    g->adj.clear();
    for (int a = 0; a < n; a++)
    {
        for (int k = g->ring_off[a]; k < g->ring_off[a + 1]; k++)
        {
            const int b = g->ring[k];
            g->adj.push_back(b);
            const int c = g->ring[k + 1 == g->ring_off[a + 1] ? g->ring_off[a] : k + 1];
            const P2 v0 = { P[b].x - P[a].x, P[b].y - P[a].y }, v1 = { P[c].x - P[a].x, P[c].y - P[a].y };
            if (v1.x * v0.y > v0.x * v1.y) continue;            // b, c do not turn the right way: the graph's boundary
            if (g->prev(b, a) != c) continue;                   // a, b, c are not a triangle
            const int d = g->prev(b, c);
            const P2 vm = { P[d].x - P[a].x, P[d].y - P[a].y };
            if (v1.x * vm.y > vm.x * v1.y) continue;            // the in-between cell must lie between b and c
            if (vm.x * v0.y > v0.x * vm.y) continue;
            g->adj.push_back(d);
        }
        g->adj_off[a + 1] = (int)g->adj.size();
    }
    g->adj_dx.resize(g->adj.size()); g->adj_dy.resize(g->adj.size()); g->adj_len.resize(g->adj.size());
    for (int a = 0; a < n; a++)
        for (int k = g->adj_off[a]; k < g->adj_off[a + 1]; k++)
        {
            g->adj_dx[k] = (double)(P[g->adj[k]].x - P[a].x); g->adj_dy[k] = (double)(P[g->adj[k]].y - P[a].y);
            g->adj_len[k] = hypot(g->adj_dx[k], g->adj_dy[k]);
        }
    return true;
}

// ---- sequences (find_grid.cc:160-343) ----
// The walk's tests on direction (cos of the turn >= 0.984) and length ratio (0.7 .. 1.4) involve only the previous
// step and the candidate step, both edges of the static graph, so they are evaluated once per pair of consecutive
// steps here instead of once per visit (every cell starts a walk towards every neighbour: ~50 visits per pair on
// a 14x14 board); the test against the running mean ratio stays in step().
void build_continuations(Graph* g)
{
    const int n = (int)g->pts.size(), E = (int)g->adj.size();
    g->cont_off.assign(E + 1, 0);
    g->cont.clear(); g->cont_ratio.clear();
    g->cont.reserve(E); g->cont_ratio.reserve(E);
    for (int a = 0; a < n; a++)
        for (int e = g->adj_off[a]; e < g->adj_off[a + 1]; e++)
        {
            const int b = g->adj[e];
            const double lx = g->adj_dx[e], ly = g->adj_dy[e], last_len = g->adj_len[e];
            for (int k = g->adj_off[b]; k < g->adj_off[b + 1]; k++)
            {
                const double dx = g->adj_dx[k], dy = g->adj_dy[k], len = g->adj_len[k];
                const double dot = lx * dx + ly * dy, prod = last_len * len;
                // far below the threshold: no need for the division (it is made whenever the outcome could be close;
                // prod == 0 gives the reference's NaN, which passes its test)
                if (prod > 0.0 && dot < 0.98 * prod) continue;
                const double cos_err = dot / prod;
                if (cos_err < kMinCos) continue;
                const double ratio = len / last_len;
                if (ratio < kMinRatio || ratio > kMaxRatio) continue;
                g->cont.push_back(k); g->cont_ratio.push_back(ratio);
            }
            g->cont_off[e + 1] = (int)g->cont.size();
        }
}

struct Walk
{
    int    e;               // most recent step (entry of Graph::adj)
    double ratio_sum;
    int    ratio_n;
};

// the first step that continues the sequence after step w->e: its entry (the cell is g.adj[entry]), or -1
int step(const Graph& g, Walk* w)
{
    for (int t = g.cont_off[w->e]; t < g.cont_off[w->e + 1]; t++)
    {
        const double ratio = g.cont_ratio[t];
        if (w->ratio_n > 2)
        {
            const double dev = ratio - w->ratio_sum / (double)w->ratio_n;
            if (dev < -kMaxRatioDeviation || dev > kMaxRatioDeviation) continue;
        }
        w->ratio_sum += ratio;
        w->ratio_n++;
        w->e = g.cont[t];
        return w->e;
    }
    return -1;
}

struct Sequence
{
    int    c0, c1, clast;
    int    e0;              // the step c0 -> c1 (entry of Graph::adj)
    double mean_dx, mean_dy;
};

// cells c0, c1 and the gridn-2 that follow (valid for a Sequence that was found: the walk is deterministic)
void sequence_cells(const Graph& g, const Sequence& s, int gridn, int* cells)
{
    cells[0] = s.c0; cells[1] = s.c1;
    Walk w = { s.e0, 0.0, 0 };
    for (int i = 0; i < gridn - 2; i++) cells[2 + i] = g.adj[step(g, &w)];
}

// find_grid.cc:776-822 (float32 throughout)
bool is_crossing(const std::vector<P2>& P, int a0, int a1, int b0, int b1)
{
    const float l0x = (float)(int)(P[a1].x - P[a0].x), l0y = (float)(int)(P[a1].y - P[a0].y);
    const float p0x = (float)(int)(P[b0].x - P[a0].x), p0y = (float)(int)(P[b0].y - P[a0].y);
    const float p1x = (float)(int)(P[b1].x - P[a0].x), p1y = (float)(int)(P[b1].y - P[a0].y);
    const float d2 = l0x * l0x + l0y * l0y;
    const float r0x = p0x * l0x + p0y * l0y, r0y = -p0x * l0y + p0y * l0x;
    const float r1x = p1x * l0x + p1y * l0y, r1y = -p1x * l0y + p1y * l0x;
    if (r0y * r1y > 0) return false;
    if ((r0x < 0 && r1x < 0) || (r0x > d2 && r1x > d2)) return false;
    const float k = r0y / (r0y - r1y);
    const float x = r0x + k * (r1x - r0x);
    return x >= 0.0f && x <= d2;
}

struct Cycle { int e[4]; };

struct Finder
{
    const Graph&             g;
    int                      gridn;
    bool                     debug;
    std::vector<Sequence>    seq;
    std::vector<int>         outer;                        // indices into seq
    std::map<int, std::vector<int>> outer_from;            // first cell -> indices into outer

    int first(int i) const { return seq[outer[i]].c0; }
    int last (int i) const { return seq[outer[i]].clast; }

    // find_grid.cc:826-960: extend e[0..count-1] to the unique 4-cycle that returns to `start`
    bool extend_cycle(Cycle* cyc, int count, int start) const
    {
        bool  found = false;
        Cycle best = {};
        const int cur = cyc->e[count - 1];
        std::map<int, std::vector<int>>::const_iterator it = outer_from.find(last(cur));
        if (it == outer_from.end()) { if (debug) fprintf(stderr, "No opposing outer edge\n"); return false; }
        const std::vector<int>& nxt = it->second;
        for (size_t k = 0; k < nxt.size(); k++)
        {
            const int e = nxt[k];
            if (last(e) == first(cur)) continue;                   // straight back
            if (count != 3)
            {
                if (last(e) == start) continue;                    // closes too early
                if (count == 2 && is_crossing(g.pts, first(cyc->e[0]), last(cyc->e[0]), first(e), last(e))) continue;
                cyc->e[count] = e;
                if (!extend_cycle(cyc, count + 1, start)) continue;
                if (found) { if (debug) fprintf(stderr, "Found non-unique 4-cycle\n"); return false; }    // two different cycles: ambiguous
                found = true;
                best = *cyc;
            }
            else
            {
                if (last(e) != start) continue;
                if (is_crossing(g.pts, first(cyc->e[1]), last(cyc->e[1]), first(e), last(e))) return false;
                cyc->e[3] = e;
                return true;
            }
        }
        if (!found) return false;
        *cyc = best;
        return true;
    }

    // find_grid.cc:962-1013
    bool opposite(const Cycle& a, const Cycle& b) const
    {
        int ia = 0, ib = -1;
        for (int k = 0; k < 4; k++) if (last(b.e[k]) == first(a.e[0])) { ib = k; break; }
        if (ib < 0)
        {
            if (debug) fprintf(stderr, "Given outer cycles are NOT equal and opposite: couldn't find a corresponding point in the two cycles\n");
            return false;
        }
        for (int k = 0; k < 4; k++)
        {
            if (first(a.e[ia]) != last(b.e[ib]) || last(a.e[ia]) != first(b.e[ib]))
            {
                if (debug) fprintf(stderr, "Given outer cycles are NOT equal and opposite\n");
                return false;
            }
            ia = (ia + 1) % 4; ib = (ib + 3) % 4;
        }
        return true;
    }

    // find_grid.cc:1015-1187: 0/1 = which cycle runs clockwise, top[] = the top edge of each; <0: give up
    int orient_cycles(const Cycle* cyc[2], int top[2]) const
    {
        const std::vector<P2>& P = g.pts;
        int v[4][2];
        for (int i = 0; i < 4; i++)
        {
            v[i][0] = (int)(P[last(cyc[0]->e[i])].x - P[first(cyc[0]->e[i])].x) / kScalePow2;
            v[i][1] = (int)(P[last(cyc[0]->e[i])].y - P[first(cyc[0]->e[i])].y) / kScalePow2;
        }
        bool sign[4];
        for (int i = 0; i < 4; i++) { const int j = (i + 1) % 4; sign[i] = (i64)v[j][0] * v[i][1] < (i64)v[i][0] * v[j][1]; }
        int clockwise;
        if      ( sign[0] &&  sign[1] &&  sign[2] &&  sign[3]) clockwise = 0;
        else if (!sign[0] && !sign[1] && !sign[2] && !sign[3]) clockwise = 1;
        else { if (debug) fprintf(stderr, "The outer edge cycles aren't convex!\n"); return -1; }

        for (int ic = 0; ic < 2; ic++)
        {
            // the two edges that share the vertex with the smallest y; the more horizontal one is the top
            i64 ymin[2] = { INT_MAX, INT_MAX };
            int edge[2] = { -1, -1 }, lo[2] = { 0, 0 }, hi[2] = { 0, 0 };
            for (int i = 0; i < 4; i++)
            {
                const int p0 = first(cyc[ic]->e[i]), p1 = last(cyc[ic]->e[i]);
                const bool up = P[p0].y < P[p1].y;
                const i64 y = up ? P[p0].y : P[p1].y;
                const int l = up ? p0 : p1, h = up ? p1 : p0;
                if (y < ymin[0])
                {
                    ymin[1] = ymin[0]; edge[1] = edge[0]; lo[1] = lo[0]; hi[1] = hi[0];
                    ymin[0] = y; edge[0] = i; lo[0] = l; hi[0] = h;
                }
                else if (y < ymin[1]) { ymin[1] = y; edge[1] = i; lo[1] = l; hi[1] = h; }
            }
            i64 v0y = (int)(P[hi[0]].y - P[lo[0]].y) / kScalePow2, v0x = (int)(P[hi[0]].x - P[lo[0]].x) / kScalePow2;
            i64 v1y = (int)(P[hi[1]].y - P[lo[1]].y) / kScalePow2, v1x = (int)(P[hi[1]].x - P[lo[1]].x) / kScalePow2;
            if (v0x < 0) v0x = -v0x;
            if (v1x < 0) v1x = -v1x;
            const i64 cross = (v0x * v1y - v0y * v1x) * (v0x * v1y - v0y * v1x);
            const i64 denom = (v0x * v0x + v0y * v0y) * (v1x * v1x + v1y * v1y);
            if ((cross < 0 ? -cross : cross) * 8 < denom * 1)                   // too close to call (sin^2 < 1/8)
            {
                if (debug)
                {
                    fprintf(stderr, "Highest 2 edges have a similar orientation. I can't tell clearly which is the more horizontal one\n");
                    const char* which[2] = { "Highest", "Second-highest" };
                    for (int q = 0; q < 2; q++)
                        fprintf(stderr, "  %s edge: (%.2f,%.2f) - (%.2f,%.2f). Highest vertex: (%.2f,%.2f)\n", which[q],
                                (double)P[first(cyc[ic]->e[edge[q]])].x / (double)kScale, (double)P[first(cyc[ic]->e[edge[q]])].y / (double)kScale,
                                (double)P[last(cyc[ic]->e[edge[q]])].x / (double)kScale,  (double)P[last(cyc[ic]->e[edge[q]])].y / (double)kScale,
                                (double)P[lo[q]].x / (double)kScale, (double)P[lo[q]].y / (double)kScale);
                    fprintf(stderr, "  sin(angle difference) as computed here: %f. Threshold: %f\n",
                            sqrt((double)(cross < 0 ? -cross : cross) / (double)denom), sqrt(1.0 / 8.0));
                }
                return -1;
            }
            const i64 l = v0y * v1x, r = v1y * v0x;
            top[ic] = (l < 0 ? -l : l) < (r < 0 ? -r : r) ? edge[0] : edge[1];
        }
        return clockwise;
    }
};

// ---- the reference's --debug artefacts (find_grid.cc:387-480, 609-778) ----
void make_executable(const char* fn) { chmod(fn, S_IRUSR | S_IRGRP | S_IROTH | S_IWUSR | S_IWGRP | S_IXUSR | S_IXGRP | S_IXOTH); }

void dump_voronoi(const Graph& g)
{
    const char* fn = "/tmp/mrgingham-2-voronoi.vnl";
    FILE* fp = fopen(fn, "w");
    if (!fp) { fprintf(stderr, "Couldn't open %s for writing\n", fn); return; }
    fprintf(fp, "#!/usr/bin/feedgnuplot --domain --dataid --with 'lines linecolor 0' --square --maxcurves 100000 --set 'yrange [:] rev'\n");
    fprintf(fp, "# x id_edge y\n");
    int i_edge = 0;
    for (size_t si = 0; si < g.sites.size(); si++)
    {
        const int a = g.sites[si];
        for (int k = g.ring_off[a]; k < g.ring_off[a + 1]; k++, i_edge++)
        {
            const int b = g.ring[k];
            fprintf(fp, "%f %d %f\n", g.pts[a].x / (double)kScale, i_edge, g.pts[a].y / (double)kScale);
            fprintf(fp, "%f %d %f\n", g.pts[b].x / (double)kScale, i_edge, g.pts[b].y / (double)kScale);
        }
    }
    fclose(fp);
    make_executable(fn);
    fprintf(stderr, "Wrote self-plotting voronoi diagram to %s\n", fn);
}

// the gridn points of a sequence, one line each: the step to the next point, dashes after the last (find_grid.cc:425-480)
void dump_intervals(FILE* fp, const Graph& g, const Sequence& s, int i_candidate, int gridn)
{
    std::vector<int> cells(gridn);
    sequence_cells(g, s, gridn, cells.data());
    for (int i = 0; i < gridn; i++)
    {
        const P2& p0 = g.pts[cells[i]];
        if (i == gridn - 1)
        {
            fprintf(fp, "%d %d %f %f - - - - - -\n", i_candidate, i, (double)p0.x / (double)kScale, (double)p0.y / (double)kScale);
            break;
        }
        const P2& p1 = g.pts[cells[i + 1]];
        const double dx = (double)(int)(p1.x - p0.x) / (double)kScale, dy = (double)(int)(p1.y - p0.y) / (double)kScale;
        fprintf(fp, "%d %d %f %f %f %f %f %f %f %f\n", i_candidate, i,
                (double)p0.x / (double)kScale, (double)p0.y / (double)kScale, (double)p1.x / (double)kScale, (double)p1.y / (double)kScale,
                dx, dy, hypot(dx, dy), atan2(dy, dx) * 180.0 / M_PI);
    }
}

// which: indices into seq (all of them when null)
void dump_candidates(const char* basename, const Graph& g, const std::vector<Sequence>& seq, const std::vector<int>* which, int gridn)
{
    const std::string sparse = std::string(basename) + ".vnl", dense = std::string(basename) + "-detailed.vnl";
    const int N = which ? (int)which->size() : (int)seq.size();
    FILE* fp = fopen(sparse.c_str(), "w");
    if (!fp) { fprintf(stderr, "Couldn't open %s for writing\n", sparse.c_str()); return; }
    fprintf(fp, "#!/usr/bin/feedgnuplot --dom --aut --square --rangesizea 3 --w 'vec size screen 0.01,20 fixed fill' --set 'yr [:] rev'\n");
    fprintf(fp, "# fromx fromy deltax deltay\n");
    for (int i = 0; i < N; i++)
    {
        const Sequence& s = seq[which ? (*which)[i] : i];
        fprintf(fp, "%f %f %f %f\n", (double)g.pts[s.c0].x / (double)kScale, (double)g.pts[s.c0].y / (double)kScale,
                s.mean_dx / (double)kScale, s.mean_dy / (double)kScale);
    }
    fclose(fp);
    make_executable(sparse.c_str());
    fprintf(stderr, "Wrote self-plotting sequence-candidate dump to %s\n", sparse.c_str());
    fp = fopen(dense.c_str(), "w");
    if (!fp) { fprintf(stderr, "Couldn't open %s for writing\n", dense.c_str()); return; }
    fprintf(fp, "# candidateid pointid fromx fromy tox toy deltax deltay len angle\n");
    for (int i = 0; i < N; i++) dump_intervals(fp, g, seq[which ? (*which)[i] : i], i, gridn);
    fclose(fp);
    fprintf(stderr, "Wrote detailed sequence-candidate dump to %s\n", dense.c_str());
}

// label(cycle, edge): the text between x and y (the cycle number, or "clockwise-top" ...)
template <class L>
void dump_cycles(const char* fn, const Finder& F, const Cycle* const* cyc, int ncyc, L label)
{
    FILE* fp = fopen(fn, "w");
    if (!fp) { fprintf(stderr, "Couldn't open %s for writing\n", fn); return; }
    fprintf(fp, "#!/usr/bin/feedgnuplot --datai --dom --aut --square --rangesizea 3 --w 'vec size screen 0.01,20 fixed fill' --set 'yr [:] rev'\n");
    fprintf(fp, "# fromx type fromy deltax deltay\n");
    for (int c = 0; c < ncyc; c++)
        for (int e = 0; e < 4; e++)
        {
            const Sequence& s = F.seq[F.outer[cyc[c]->e[e]]];
            fprintf(fp, "%f %s %f %f %f\n", (double)F.g.pts[s.c0].x / (double)kScale, label(c, e).c_str(), (double)F.g.pts[s.c0].y / (double)kScale,
                    s.mean_dx / (double)kScale, s.mean_dy / (double)kScale);
        }
    fclose(fp);
    make_executable(fn);
    fprintf(stderr, "Wrote outer edge cycle dump to %s\n", fn);
}

// The walk of step() from one cell with the reference's commentary on stderr (find_grid.cc:216-310): every neighbour
// considered, why it is rejected, which one is accepted. Same tests in the same order as build_continuations() + step().
int step_traced(const Graph& g, Walk* w)
{
    const int a_e = w->e, b = g.adj[a_e];
    const double lx = g.adj_dx[a_e], ly = g.adj_dy[a_e], last_len = g.adj_len[a_e];
    for (int k = g.adj_off[b]; k < g.adj_off[b + 1]; k++)
    {
        const int c = g.adj[k];
        fprintf(stderr, "Considering connection in sequence from (%d,%d) -> (%d,%d); delta (%d,%d) ..... \n",
                (int)g.pts[b].x / kScale, (int)g.pts[b].y / kScale, (int)g.pts[c].x / kScale, (int)g.pts[c].y / kScale,
                (int)(g.pts[c].x - g.pts[b].x) / kScale, (int)(g.pts[c].y - g.pts[b].y) / kScale);
        const double dx = g.adj_dx[k], dy = g.adj_dy[k], len = g.adj_len[k];
        const double cos_err = (lx * dx + ly * dy) / (last_len * len);
        if (cos_err < kMinCos)
        {
            fprintf(stderr, "..... rejecting. Angle is wrong. I wanted cos_err>=threshold, but saw %f<%f\n", cos_err, kMinCos);
            continue;
        }
        const double ratio = len / last_len;
        if (ratio < kMinRatio || ratio > kMaxRatio)
        {
            fprintf(stderr, "..... rejecting. Lengths are wrong. I wanted abs(length_ratio)<=threshold, but saw %f<%f or %f>%f\n",
                    ratio, kMinRatio, ratio, kMaxRatio);
            continue;
        }
        if (w->ratio_n > 2)
        {
            const double dev = ratio - w->ratio_sum / (double)w->ratio_n;
            if (dev < -kMaxRatioDeviation || dev > kMaxRatioDeviation)
            {
                fprintf(stderr, "..... rejecting. Lengths are wrong. I wanted abs(length_ratio_deviation)<=threshold, but saw %f>%f\n",
                        fabs(dev), kMaxRatioDeviation);
                continue;
            }
        }
        w->ratio_sum += ratio;
        w->ratio_n++;
        w->e = k;
        fprintf(stderr, "..... accepting!\n\n");
        return k;
    }
    return -1;
}
}   // namespace

int voronoi_neighbours(const int* xy, int npoints, int* ring_off, int* ring, int ring_cap)
{
    Graph g;
    if (npoints <= 0 || !build_graph(&g, xy, npoints)) return -1;
    for (int i = 0; i <= npoints; i++) ring_off[i] = g.ring_off[i];
    for (int k = 0; k < g.ring_off[npoints] && k < ring_cap; k++) ring[k] = g.ring[k];
    return g.ring_off[npoints];
}

bool find_grid_from_points(const int* xy, int npoints, int gridn, double* xy_out, const GridDebug* dbg)
{
    if (npoints <= 0 || gridn < 2 || !xy || !xy_out) return false;
    const bool debug = dbg && dbg->dump;
    Graph g;
    if (!build_graph(&g, xy, npoints)) return false;
    build_continuations(&g);
    if (debug) dump_voronoi(g);
    Finder F = { g, gridn, debug, {}, {}, {} };

    // debug_sequence: the cell nearest to the given pixel is the one whose walks are narrated (find_grid.cc:515-540)
    int traced = -1;
    if (dbg && dbg->sequence)
    {
        unsigned long best = (unsigned long)(-1L);
        for (size_t si = 0; si < g.sites.size(); si++)
        {
            const int c = g.sites[si];
            const long dx = (long)(g.pts[c].x - (i64)kScale * dbg->seq_x), dy = (long)(g.pts[c].y - (i64)kScale * dbg->seq_y);
            const unsigned long d2 = (unsigned long)(dx * dx + dy * dy);
            if (d2 < best) { best = d2; traced = c; }
        }
        fprintf(stderr, "============== Looking at sequences from (%d,%d)\n", (int)g.pts[traced].x / kScale, (int)g.pts[traced].y / kScale);
    }

    // every run of gridn cells, from every cell towards every neighbour (find_grid.cc:505-566)
    for (size_t si = 0; si < g.sites.size(); si++)
    {
        const int c = g.sites[si];
        for (int k = g.adj_off[c]; k < g.adj_off[c + 1]; k++)
        {
            const int c1 = g.adj[k];
            if (c == traced)
                fprintf(stderr, "\n\n====== Looking at adjacent point (%d,%d)\n", (int)g.pts[c1].x / kScale, (int)g.pts[c1].y / kScale);
            Walk w = { k, 0.0, 0 };
            double mx = g.adj_dx[k], my = g.adj_dy[k];
            int clast = -1;
            for (int i = 0; i < gridn - 2; i++)
            {
                const int e = c == traced ? step_traced(g, &w) : step(g, &w);
                if (e < 0) { clast = -1; break; }
                mx += g.adj_dx[e]; my += g.adj_dy[e];
                clast = g.adj[e];
            }
            if (clast < 0) continue;
            const Sequence s = { c, c1, clast, k, mx / (double)(gridn - 1), my / (double)(gridn - 1) };
            F.seq.push_back(s);
        }
    }
    if (debug)
    {
        dump_candidates("/tmp/mrgingham-3-candidates", g, F.seq, nullptr, gridn);
        fprintf(stderr, "got %zd points\n", (size_t)npoints);
        fprintf(stderr, "got %zd sequence candidates\n", F.seq.size());
    }

    // the board's outer edges start at cells that start at least two sequences (find_grid.cc:1244-1275)
    std::map<int, int> started;
    for (size_t i = 0; i < F.seq.size(); i++) started[F.seq[i].c0]++;
    for (size_t i = 0; i < F.seq.size(); i++) if (started[F.seq[i].c0] >= 2) F.outer.push_back((int)i);
    if (F.outer.size() < 8)
    {
        if (debug) fprintf(stderr, "Too few candidates for an outer edge of the grid. Needed at least 8, got %d\n", (int)F.outer.size());
        return false;
    }
    if (debug) dump_candidates("/tmp/mrgingham-4-outer-edges", g, F.seq, &F.outer, gridn);
    for (size_t i = 0; i < F.outer.size(); i++) F.outer_from[F.first((int)i)].push_back((int)i);

    // 4-cycles of outer edges (find_grid.cc:1290-1317)
    std::vector<Cycle> cycles;
    std::set<int> used;
    for (int i = 0; i < (int)F.outer.size(); i++)
    {
        if (used.count(i)) continue;
        Cycle c = {}; c.e[0] = i;
        if (!F.extend_cycle(&c, 1, F.first(i))) continue;
        cycles.push_back(c);
        for (int k = 0; k < 4; k++) used.insert(c.e[k]);
    }
    if (debug && !cycles.empty())
    {
        std::vector<const Cycle*> all(cycles.size());
        for (size_t i = 0; i < cycles.size(); i++) all[i] = &cycles[i];
        dump_cycles("/tmp/mrgingham-5-outer-edge-cycles", F, all.data(), (int)all.size(), [](int c, int) { return std::to_string(c); });
    }
    if (cycles.size() < 2)
    {
        if (debug) fprintf(stderr, "Found too few 4-cycles. Needed at least 2, got %d\n", (int)cycles.size());
        return false;
    }

    // exactly one pair of cycles running the same way round in opposite directions (find_grid.cc:1329-1353)
    int pair[2] = { -1, -1 };
    for (size_t a = 0; a < cycles.size(); a++)
        for (size_t b = a + 1; b < cycles.size(); b++)
            if (F.opposite(cycles[a], cycles[b]))
            {
                if (pair[0] >= 0)
                {
                    if (debug) fprintf(stderr, "Found more than one equal-and-opposite pair of outer-edge cycles. Giving up\n");
                    return false;
                }
                pair[0] = (int)a; pair[1] = (int)b;
            }
    if (pair[0] < 0)
    {
        if (debug) fprintf(stderr, "Didn't find any equal-and-opposite pairs of outer-edge cycles. Giving up\n");
        return false;
    }

    const Cycle* cyc[2] = { &cycles[pair[0]], &cycles[pair[1]] };
    int top[2];
    const int cw = F.orient_cycles(cyc, top);
    if (cw < 0) return false;
    if (debug)
        dump_cycles("/tmp/mrgingham-6-identified-outer-edge-cycle", F, cyc, 2, [&](int c, int e)
                    { return std::string(c == cw ? "clockwise" : "counterclockwise") + (top[c] == e ? "-top" : ""); });

    // rows run from the i-th cell of the left edge to the i-th cell of the right edge (find_grid.cc:1385-1433)
    std::map<int, std::vector<int>> seq_from;
    for (size_t i = 0; i < F.seq.size(); i++) seq_from[F.seq[i].c0].push_back((int)i);
    auto from_to = [&](int a, int b) -> int
    {
        std::map<int, std::vector<int>>::const_iterator it = seq_from.find(a);
        if (it == seq_from.end()) return -1;
        for (size_t k = 0; k < it->second.size(); k++) if (F.seq[it->second[k]].clast == b) return it->second[k];
        return -1;
    };
    std::vector<int> rows(gridn), left(gridn), right(gridn), cells(gridn);
    rows[0] = F.outer[cyc[cw]->e[top[cw]]];
    sequence_cells(g, F.seq[F.outer[cyc[1 - cw]->e[(top[1 - cw] + 1) % 4]]], gridn, left.data());
    sequence_cells(g, F.seq[F.outer[cyc[cw]->e[(top[cw] + 1) % 4]]], gridn, right.data());
    for (int i = 1; i < gridn; i++)
    {
        rows[i] = from_to(left[i], right[i]);
        if (rows[i] < 0) { if (debug) fprintf(stderr, "Couldn't find sequence in row %d\n", i); return false; }
        if (from_to(right[i], left[i]) < 0)
        {
            if (debug) fprintf(stderr, "Row %d: left-to-right sequence was found, but right-to-left sequence doesn't exist!\n", i);
            return false;
        }
    }
    for (int i = 0; i < gridn; i++)
    {
        sequence_cells(g, F.seq[rows[i]], gridn, cells.data());
        for (int k = 0; k < gridn; k++)
        {
            xy_out[2 * (i * gridn + k)]     = (double)g.pts[cells[k]].x / (double)kScale;
            xy_out[2 * (i * gridn + k) + 1] = (double)g.pts[cells[k]].y / (double)kScale;
        }
    }
    if (debug) fprintf(stderr, "Success. Found grid\n");
    return true;
}

}
