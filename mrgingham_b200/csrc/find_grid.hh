// Host-side grid finder (find_grid.cu): mrgingham::find_grid_from_points, find_grid.cc:1216-1445.
#pragma once

namespace mrgb200
{
// xy: npoints (x,y) pairs scaled by 1000 (PointInt); xy_out: gridn*gridn (x,y) doubles in pixels, rows from the
// board's top edge. Returns true iff exactly one gridn x gridn grid was found (xy_out untouched otherwise).
// The neighbour graph the grid finder works on (what the reference's --debug voronoi dump shows, find_grid.cc:
// 391-430): for point i, ring[ring_off[i] .. ring_off[i+1]) are the points whose Voronoi cells share an edge
// with its cell, counter-clockwise in (x,y) from the +x direction. Returns the total ring length (which may
// exceed ring_cap; only ring_cap entries are written) or <0.
int voronoi_neighbours(const int* xy, int npoints, int* ring_off, int* ring, int ring_cap);

// The reference's diagnostics (find_grid.cc:387-423, 425-480, 609-778 and the fprintf(stderr) of :216-343, :505-566,
// :1216-1445): dump = its `debug` (the self-plotting /tmp/mrgingham-2-voronoi.vnl, -3-candidates[-detailed].vnl,
// -4-outer-edges[-detailed].vnl, -5-outer-edge-cycles, -6-identified-outer-edge-cycle and the messages that say why a
// grid was not found); sequence = its debug_sequence (a trace on stderr of every connection considered from the point
// nearest to (seq_x, seq_y), in pixels).
struct GridDebug { bool dump; bool sequence; int seq_x, seq_y; };

bool find_grid_from_points(const int* xy, int npoints, int gridn, double* xy_out, const GridDebug* debug = nullptr);
}
