// TEST INFRASTRUCTURE: runs the border-state formulation of mrgingham_b200/csrc/blob_walk.cuh (the code the CUDA
// kernels of blobs.cu are built from) on the CPU, so that tests/test_blob_walk_host.py can compare it with the
// oracle's Suzuki-Abe restatement without a GPU. Built by the test with g++.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include <algorithm>
#include "../../mrgingham_b200/csrc/blob_walk.cuh"

using namespace mrgb200::blobwalk;

struct Found { int pos, x, y, k, n; long long a00; };

// bin: h x w bytes (nonzero = foreground). Output as blob_oracle_find_contours: contours in REVERSE discovery
// order, xy = [total][2], lens = [ncont], area2 = [ncont] (the a00 sums). Returns the number of contours
// (isolated pixels included as one-point contours), or -1 if a capacity is too small.
extern "C" int blob_walk_host_contours(const uint8_t* bin, int w, int h, int stride, int32_t* xy, int max_pts,
                                       int32_t* lens, long long* area2, int max_cont)
{
    const int wpr = plane_wpr(w), words = (w + 31) / 32;
    std::vector<uint32_t> storage((size_t)wpr * plane_rows(h), 0u);
    uint32_t* plane = storage.data() + plane_origin(w);
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++)
            if (bin[(size_t)y * stride + x]) plane[(size_t)y * wpr + (x >> 5)] |= 1u << (x & 31);
    PlaneRef P{ plane, w, h, wpr };
    BitWindow F, Bk;
    F.init(); Bk.init();
    std::vector<Found> found;
    auto word = [&](int y, int wd) -> uint32_t { return plane[y * wpr + wd]; };       // y in [-1, h], wd in [-1, words]
    for (int y = 0; y < h; y++)
        for (int wd = 0; wd < words; wd++)
        {
            uint32_t outer, hole;
            candidate_masks(word(y, wd), word(y, wd - 1), word(y - 1, wd), word(y - 1, wd - 1), word(y - 1, wd + 1), &outer, &hole);
            for (int kind = 0; kind < 2; kind++)
            {
                uint32_t m = kind ? hole : outer;
                while (m)
                {
                    const int b = ffs32(m) - 1; m &= m - 1;
                    const int x = wd * 32 + b, pos = y * w + x;
                    int sx = x, sy = y, sk = 1;
                    if (kind == 0)
                    {
                        if (!outer_start(P, F, x, y, &sk)) { found.push_back(Found{ pos, x, y, -1, 1, 0 }); continue; }
                    }
                    else sx = x - 1;
                    int n; long long a00;
                    if (verify_start(P, F, Bk, sx, sy, sk, pos, &n, &a00, 4LL * w * h + 16))
                        found.push_back(Found{ pos, sx, sy, sk, n, a00 });
                }
            }
        }
    std::sort(found.begin(), found.end(), [](const Found& a, const Found& b) { return a.pos > b.pos; });
    if ((int)found.size() > max_cont) return -1;
    int o = 0, c = 0;
    for (const Found& f : found)
    {
        if (o + f.n > max_pts) return -1;
        lens[c] = f.n; area2[c] = f.a00; c++;
        int x = f.x, y = f.y, k = f.k, disc;
        for (int i = 0; i < f.n; i++)
        {
            xy[2*o] = x; xy[2*o + 1] = y; o++;
            if (f.k >= 0) step_fwd(P, F, x, y, k, &disc);
        }
    }
    return (int)found.size();
}
