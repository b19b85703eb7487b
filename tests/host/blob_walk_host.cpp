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

// The same contours through the SEGMENT formulation (blob_walk.cuh, "borders in SEGMENTS"): every segment start --
// candidates, row cuts, column cuts -- is walked forward to the next one; candidates then follow the chain of segments.
#include <unordered_map>
struct Seg { unsigned long long start, end; int n, min_disc, pos, x, y, k; long long a00; bool cand; };

extern "C" int blob_walk_host_contours_segments(const uint8_t* bin, int w, int h, int stride, int32_t* xy, int max_pts,
                                                int32_t* lens, long long* area2, int max_cont, int* nsegments)
{
    const int wpr = plane_wpr(w), words = (w + 31) / 32;
    std::vector<uint32_t> storage((size_t)wpr * plane_rows(h), 0u);
    uint32_t* plane = storage.data() + plane_origin(w);
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++)
            if (bin[(size_t)y * stride + x]) plane[(size_t)y * wpr + (x >> 5)] |= 1u << (x & 31);
    PlaneRef P{ plane, w, h, wpr };
    BitWindow F; F.init();
    std::vector<Seg> segs;
    std::unordered_map<unsigned long long, int> index;
    std::vector<Found> found;
    auto word = [&](int y, int wd) -> uint32_t { return plane[y * wpr + wd]; };
    auto add_start = [&](int x, int y, int k, bool cand, int pos)
    {
        const unsigned long long key = state_key(0, x, y, k);
        auto it = index.find(key);
        if (it != index.end()) return;                           // the same state, queued twice
        Seg s; s.start = key; s.x = x; s.y = y; s.k = k;
        int disc, cpos; bool st;
        classify_state(P, F, x, y, k, &disc, &st, &cpos);
        if (!st) abort();                                        // scan and walker must agree on what a start is
        if (cand && cpos != pos) abort();                        // ... and on which starts are candidates, discovered where
        s.cand = cpos >= 0; s.pos = cpos;
        s.min_disc = disc < 0 ? 0x7fffffff : disc; s.n = 0; s.a00 = 0;
        int cx = x, cy = y, ck = k;
        for (;;)
        {
            const int px = cx, py = cy;
            step_fwd_ex(P, F, cx, cy, ck, &disc, &st);
            s.a00 += (long long)(px * cy - cx * py); s.n++;
            if (st) break;
            if (disc >= 0 && disc < s.min_disc) s.min_disc = disc;
        }
        s.end = state_key(0, cx, cy, ck);
        index[key] = (int)segs.size();
        segs.push_back(s);
    };
    for (int y = 0; y < h; y++)
        for (int wd = 0; wd < words; wd++)
        {
            uint32_t outer, hole, cut[4];
            candidate_masks(word(y, wd), word(y, wd - 1), word(y - 1, wd), word(y - 1, wd - 1), word(y - 1, wd + 1), &outer, &hole);
            cut_masks(word(y, wd), word(y, wd - 1), word(y, wd + 1), word(y - 1, wd), word(y + 1, wd), wd, y, &cut[0], &cut[1], &cut[2], &cut[3]);
            const int cut_dir[4] = { 4, 0, 2, 6 };
            for (int kind = 0; kind < 6; kind++)
            {
                uint32_t m = kind == 0 ? outer : kind == 1 ? hole : cut[kind - 2];
                while (m)
                {
                    const int b = ffs32(m) - 1; m &= m - 1;
                    const int x = wd * 32 + b, pos = y * w + x;
                    int k = 1;
                    if (kind == 1) { add_start(x - 1, y, 1, true, pos); continue; }
                    if (!state_after(P, F, x, y, kind == 0 ? 4 : cut_dir[kind - 2], &k))
                    {
                        if (kind == 0) found.push_back(Found{ pos, x, y, -1, 1, 0 });   // isolated pixel
                        continue;
                    }
                    add_start(x, y, k, kind == 0, pos);
                }
            }
        }
    *nsegments = (int)segs.size();
    for (const Seg& c : segs)
    {
        if (!c.cand) continue;
        long long a = 0; int n = 0; bool ok = true;
        const Seg* cur = &c;
        for (;;)
        {
            if (cur->min_disc < c.pos) { ok = false; break; }
            a += cur->a00; n += cur->n;
            if (cur->end == c.start) break;
            auto it = index.find(cur->end);
            if (it == index.end()) abort();
            cur = &segs[it->second];
        }
        if (ok) found.push_back(Found{ c.pos, c.x, c.y, c.k, n, a });
    }
    std::sort(found.begin(), found.end(), [](const Found& a, const Found& b) { return a.pos > b.pos; });
    if ((int)found.size() > max_cont) return -1;
    int o = 0, cn = 0;
    for (const Found& f : found)
    {
        if (o + f.n > max_pts) return -1;
        lens[cn] = f.n; area2[cn] = f.a00; cn++;
        int x = f.x, y = f.y, k = f.k, disc;
        for (int i = 0; i < f.n; i++)
        {
            xy[2*o] = x; xy[2*o + 1] = y; o++;
            if (f.k >= 0) step_fwd(P, F, x, y, k, &disc);
        }
    }
    return (int)found.size();
}
