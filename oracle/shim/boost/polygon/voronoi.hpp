// TEST INFRASTRUCTURE ONLY -- never shipped, never included by the product library.
//
// Stand-in for the part of Boost.Polygon's Voronoi API that the reference's grid finder reads
// (/root/reference/find_grid.cc:7,17-36,86-140,399-411,522-543,1226-1227), so that the UNMODIFIED
// find_grid.cc / mrgingham.cc compile here (Boost is not in this image) and serve as the checker
// of the library's own grid finder (oracle/Makefile, target `refgrid`).
//
// What the reference uses and what this header therefore provides:
//   boost::polygon::voronoi_diagram<double>   cells() -> container of cell_type, in creation order
//   cell_type    source_index(), incident_edge()
//   edge_type    cell(), twin(), next(), prev()
//   construct_voronoi(first, last, &diagram)  over a point type described by point_traits<>
//   geometry_concept<> / point_traits<> / point_concept / orientation_2d (HORIZONTAL) for the
//   traits specialisations at find_grid.cc:18-31
//
// What is modelled (the same three conventions DESIGN.md 5d and oracle/grid_oracle.py state):
//   * two sites are neighbours iff their Voronoi cells share an edge of NON-ZERO length; decided
//     here from the definition -- the centres on the bisector of a and b whose circle through a, b
//     keeps every other site strictly outside form an interval of non-zero length -- in exact
//     128-bit integer arithmetic, O(n^3). Boost's voronoi_diagram drops zero-length edges between
//     cocircular sites too.
//   * repeated points are one site (Boost sorts and uniques its site events); cells are created
//     in sorted-site order (x, then y), which is the order Boost's sweep processes site events in.
//   * next() walks a cell's edges counter-clockwise in (x, y) ("clockwise" in an image with y
//     down, find_grid.cc:40-41); the ring closes over the unbounded side of hull cells as Boost's
//     does. incident_edge(): Boost's choice is an artefact of its sweep; here it is the edge to the
//     first neighbour counter-clockwise from the +x direction. A site without neighbours has
//     incident_edge() == NULL.
#pragma once

#include <stddef.h>
#include <math.h>      // the real header pulls <cmath> in; find_grid.cc relies on that for hypot/atan2/fabs/M_PI
#include <cmath>
#include <algorithm>
#include <iterator>
#include <vector>

namespace boost { namespace polygon {

enum orientation_2d { HORIZONTAL = 0, VERTICAL = 1 };
struct point_concept {};
template <typename T> struct geometry_concept {};
template <typename T> struct point_traits {};

template <typename T>
class voronoi_diagram
{
public:
    class cell_type;
    class edge_type
    {
    public:
        const cell_type* cell() const { return cell_; }
        const edge_type* twin() const { return twin_; }
        const edge_type* next() const { return next_; }
        const edge_type* prev() const { return prev_; }
        const cell_type* cell_;
        const edge_type *twin_, *next_, *prev_;
    };
    class cell_type
    {
    public:
        size_t           source_index()  const { return source_index_; }
        const edge_type* incident_edge() const { return incident_edge_; }
        size_t           source_index_;
        const edge_type* incident_edge_;
    };
    typedef std::vector<cell_type> cell_container_type;
    typedef std::vector<edge_type> edge_container_type;

    voronoi_diagram() {}
    const cell_container_type& cells() const { return cells_; }
    const edge_container_type& edges() const { return edges_; }
    size_t num_cells() const { return cells_.size(); }

    cell_container_type cells_;
    edge_container_type edges_;

private:
    voronoi_diagram(const voronoi_diagram&);
    void operator=(const voronoi_diagram&);
};

namespace shim_detail
{
    typedef __int128 wide;
    struct site { long long x, y; size_t source; };

    // p1/q1 < p2/q2 for q1, q2 != 0
    inline bool frac_less(wide p1, wide q1, wide p2, wide q2)
    {
        if (q1 < 0) { p1 = -p1; q1 = -q1; }
        if (q2 < 0) { p2 = -p2; q2 = -q2; }
        return p1 * q2 < p2 * q1;
    }

    // Do the cells of sites ia and ib share an edge of non-zero length? A circle through a and b
    // has its centre at (a+b)/2 + t*d, d = (b-a) turned by 90 degrees; site c is strictly outside it
    // iff t*alpha_c > beta_c with the integers below. The admissible t form the open interval
    // (max over alpha>0 of beta/alpha, min over alpha<0 of beta/alpha).
    inline bool share_an_edge(const std::vector<site>& s, size_t ia, size_t ib)
    {
        const wide ax = s[ia].x, ay = s[ia].y, bx = s[ib].x, by = s[ib].y;
        const wide dx = ay - by, dy = bx - ax;
        bool have_lo = false, have_hi = false;
        wide lo_p = 0, lo_q = 1, hi_p = 0, hi_q = 1;
        for (size_t ic = 0; ic < s.size(); ic++)
        {
            if (ic == ia || ic == ib) continue;
            const wide cx = s[ic].x, cy = s[ic].y;
            const wide ex = ax - cx, ey = ay - cy;
            const wide alpha = 2 * (dx * ex + dy * ey);
            const wide beta  = (ax * ax + ay * ay) - (cx * cx + cy * cy) - ((ax + bx) * ex + (ay + by) * ey);
            if (alpha == 0)
            {
                if (beta >= 0) return false;      // c lies on the line through a and b, between or on them
                continue;
            }
            if (alpha > 0) { if (!have_lo || frac_less(lo_p, lo_q, beta, alpha)) { lo_p = beta; lo_q = alpha; have_lo = true; } }
            else           { if (!have_hi || frac_less(beta, alpha, hi_p, hi_q)) { hi_p = beta; hi_q = alpha; have_hi = true; } }
            if (have_lo && have_hi && !frac_less(lo_p, lo_q, hi_p, hi_q)) return false;
        }
        return true;
    }

    // is direction u before direction v, counter-clockwise starting at +x?
    inline bool ccw_before(long long ux, long long uy, long long vx, long long vy)
    {
        const int hu = (uy > 0 || (uy == 0 && ux > 0)) ? 0 : 1;
        const int hv = (vy > 0 || (vy == 0 && vx > 0)) ? 0 : 1;
        if (hu != hv) return hu < hv;
        return (wide)ux * vy - (wide)uy * vx > 0;
    }
}

template <typename PointIterator, typename T>
void construct_voronoi(PointIterator first, PointIterator last, voronoi_diagram<T>* vd)
{
    using namespace shim_detail;
    typedef typename std::iterator_traits<PointIterator>::value_type point_type;

    std::vector<site> all;
    size_t idx = 0;
    for (PointIterator it = first; it != last; ++it, ++idx)
    {
        site s;
        s.x = point_traits<point_type>::get(*it, HORIZONTAL);
        s.y = point_traits<point_type>::get(*it, VERTICAL);
        s.source = idx;
        all.push_back(s);
    }
    std::sort(all.begin(), all.end(), [](const site& a, const site& b)
              { return a.x != b.x ? a.x < b.x : a.y != b.y ? a.y < b.y : a.source < b.source; });
    std::vector<site> s;
    for (size_t i = 0; i < all.size(); i++)
        if (s.empty() || s.back().x != all[i].x || s.back().y != all[i].y) s.push_back(all[i]);

    const size_t n = s.size();
    std::vector<std::vector<size_t> > ring(n);
    for (size_t a = 0; a < n; a++)
        for (size_t b = a + 1; b < n; b++)
            if (share_an_edge(s, a, b)) { ring[a].push_back(b); ring[b].push_back(a); }
    size_t nedges = 0;
    std::vector<size_t> first_edge(n);
    for (size_t a = 0; a < n; a++)
    {
        std::sort(ring[a].begin(), ring[a].end(), [&](size_t p, size_t q)
                  { return ccw_before(s[p].x - s[a].x, s[p].y - s[a].y, s[q].x - s[a].x, s[q].y - s[a].y); });
        first_edge[a] = nedges;
        nedges += ring[a].size();
    }

    typedef typename voronoi_diagram<T>::cell_type cell_type;
    typedef typename voronoi_diagram<T>::edge_type edge_type;
    vd->cells_.assign(n, cell_type());
    vd->edges_.assign(nedges, edge_type());
    for (size_t a = 0; a < n; a++)
    {
        const size_t k = ring[a].size();
        cell_type& c = vd->cells_[a];
        c.source_index_  = s[a].source;
        c.incident_edge_ = k ? &vd->edges_[first_edge[a]] : NULL;
        for (size_t i = 0; i < k; i++)
        {
            edge_type& e = vd->edges_[first_edge[a] + i];
            const size_t b = ring[a][i];
            const size_t back = std::find(ring[b].begin(), ring[b].end(), a) - ring[b].begin();
            e.cell_ = &c;
            e.twin_ = &vd->edges_[first_edge[b] + back];
            e.next_ = &vd->edges_[first_edge[a] + (i + 1) % k];
            e.prev_ = &vd->edges_[first_edge[a] + (i + k - 1) % k];
        }
    }
}

}}
