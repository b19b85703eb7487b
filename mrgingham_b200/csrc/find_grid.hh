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

bool find_grid_from_points(const int* xy, int npoints, int gridn, double* xy_out);
}
