// TEST INFRASTRUCTURE ONLY -- never shipped, never linked into the product library.
//
// extern "C" handles onto the UNMODIFIED reference grid finder and board pipeline, compiled by
// oracle/Makefile (target `refgrid`) from /root/reference/find_grid.cc and /root/reference/mrgingham.cc
// where they lie, against two stand-ins in oracle/shim/: the cv::Mat shim and
// boost/polygon/voronoi.hpp (Boost is not in this image; the header states exactly what it models).
// The corner detector underneath is the reference's own find_chessboard_corners.cc + ChESS.c.
//
// Reference entry points used (file:line in /root/reference):
//   mrgingham::find_grid_from_points              find_grid.cc:1216
//   mrgingham::find_chessboard_from_image_array   mrgingham.cc:106   (level loop :127-138, refinement :57-99)
#include <vector>
#include <stdint.h>
#include <stdlib.h>
#include <stdio.h>

#include "mrgingham.hh"
#include "find_blobs.hh"
#include <boost/polygon/voronoi.hpp>

// a point type of this file's own for ref_shim_voronoi_rings (find_grid.cc specialises the traits for PointInt
// inside its own translation unit, find_grid.cc:18-31)
namespace { struct shim_pt { int x, y; }; }
namespace boost { namespace polygon {
    template <> struct geometry_concept<shim_pt> { typedef point_concept type; };
    template <> struct point_traits<shim_pt>
    {
        typedef int coordinate_type;
        static inline coordinate_type get(const shim_pt& p, orientation_2d o) { return o == HORIZONTAL ? p.x : p.y; }
    };
}}

#define API extern "C" __attribute__((visibility("default")))

// mrgingham.cc references the blob detector (find_blobs.cc wraps cv::SimpleBlobDetector, which cannot be
// built here); these stand-ins only satisfy the linker and are never reached by the calls below.
// (Not in the `hybrid` build of oracle/Makefile: there find_blobs.hh is this repo's adapter header.)
#ifndef REF_CAPI_HYBRID
namespace mrgingham
{
    bool find_blobs_from_image_array(std::vector<PointInt>*, const cv::Mat&, bool)
    {
        fprintf(stderr, "ref_grid_capi: the reference's blob detector is not part of this build\n");
        return false;
    }
    bool find_blobs_from_image_file(std::vector<PointInt>*, const char*, bool)
    {
        fprintf(stderr, "ref_grid_capi: the reference's blob detector is not part of this build\n");
        return false;
    }
}
#endif

// points: n x (x, y) ints scaled by 1000 (PointInt). out: gridn*gridn x (x, y) doubles. Returns 1 if a grid was found.
API int ref_find_grid_from_points(const int* xy, int n, int gridn, double* out)
{
    std::vector<mrgingham::PointInt> pts;
    pts.reserve(n);
    for(int i = 0; i < n; i++) pts.push_back(mrgingham::PointInt(xy[2*i], xy[2*i+1]));
    std::vector<mrgingham::PointDouble> grid;
    const bool found = mrgingham::find_grid_from_points(grid, pts, gridn);
    if(!found) return 0;
    if((int)grid.size() != gridn*gridn) return -1;
    for(int i = 0; i < gridn*gridn; i++) { out[2*i] = grid[i].x; out[2*i+1] = grid[i].y; }
    return 1;
}

// The same with the reference's diagnostics switched on: debug (the /tmp/mrgingham-{2..6}-* dumps and stderr messages,
// find_grid.cc:387-480, 609-778) and debug_sequence (the stderr trace from the point nearest to pixel (sx, sy); sx < 0: off).
API int ref_find_grid_from_points_debug(const int* xy, int n, int gridn, double* out, int debug, int sx, int sy)
{
    std::vector<mrgingham::PointInt> pts;
    pts.reserve(n);
    for(int i = 0; i < n; i++) pts.push_back(mrgingham::PointInt(xy[2*i], xy[2*i+1]));
    std::vector<mrgingham::PointDouble> grid;
    mrgingham::debug_sequence_t ds = {};
    if(sx >= 0 && sy >= 0) { ds.dodebug = true; ds.pt.x = sx; ds.pt.y = sy; }
    const bool found = mrgingham::find_grid_from_points(grid, pts, gridn, debug != 0, ds);
    fflush(stderr);
    if(!found) return 0;
    if((int)grid.size() != gridn*gridn) return -1;
    for(int i = 0; i < gridn*gridn; i++) { out[2*i] = grid[i].x; out[2*i+1] = grid[i].y; }
    return 1;
}

// The whole board pipeline on one image. level < 0: the reference's 3,2,1,0 loop. Returns the level the board
// was found at, or -1. levels_out (gridn*gridn, may be NULL when !refine): the level each point was refined to.
API int ref_find_chessboard_from_image_array(const uint8_t* image, int rows, int cols, int stride,
                                             int gridn, int level, int refine,
                                             double* out, signed char* levels_out)
{
    cv::Mat m(rows, cols, CV_8UC1, (void*)image, (size_t)stride);
    std::vector<mrgingham::PointDouble> grid;
    signed char* refinement_level = NULL;
    const int found = mrgingham::find_chessboard_from_image_array(grid, refine ? &refinement_level : NULL,
                                                                  gridn, m, level);
    if(found >= 0)
    {
        if((int)grid.size() != gridn*gridn) { free(refinement_level); return -2; }
        for(int i = 0; i < gridn*gridn; i++)
        {
            out[2*i] = grid[i].x; out[2*i+1] = grid[i].y;
            if(levels_out) levels_out[i] = refinement_level ? refinement_level[i] : (signed char)found;
        }
    }
    free(refinement_level);
    return found;
}

// The neighbour rings of the Voronoi stand-in, exposed so that tests can pin it against the definition-based
// Python oracle (oracle/grid_oracle.py): for every cell in creation order its source index, then its
// neighbours' source indices in next() order starting at incident_edge(). Layout of `out`:
// [ncells, then per cell: source, k, n_0 .. n_k-1]; returns the number of ints needed (written only if <= cap).
API int ref_shim_voronoi_rings(const int* xy, int n, int* out, int cap)
{
    std::vector<shim_pt> pts(n);
    for(int i = 0; i < n; i++) { pts[i].x = xy[2*i]; pts[i].y = xy[2*i+1]; }
    boost::polygon::voronoi_diagram<double> vd;
    construct_voronoi(pts.begin(), pts.end(), &vd);
    std::vector<int> v;
    v.push_back((int)vd.cells().size());
    for(auto it = vd.cells().begin(); it != vd.cells().end(); ++it)
    {
        v.push_back((int)it->source_index());
        const size_t at = v.size();
        v.push_back(0);
        const auto* e0 = it->incident_edge();
        if(e0 == NULL) continue;
        const auto* e = e0;
        do { v.push_back((int)e->twin()->cell()->source_index()); v[at]++; e = e->next(); } while(e != e0);
    }
    if((int)v.size() <= cap) for(size_t i = 0; i < v.size(); i++) out[i] = v[i];
    return (int)v.size();
}
