// The blob path: find_blobs_from_image_array() (find_blobs.cc:14-46), i.e. cv::SimpleBlobDetector
// with minArea=20, maxArea=80000, minDistBetweenBlobs=5, blobColor=0 (find_blobs.cc:18-22) and
// OpenCV's defaults otherwise (thresholds 50,60,..,210; minRepeatability 2; inertia >= 0.1;
// convexity >= 0.95), then PointInt((int)(pt.x*1000 + 0.5), ...) (find_blobs.cc:40-41).
//
// The algorithm is OpenCV's (third party, not under /root/reference); what is reproduced, and
// pinned by the CPU restatement used in the tests against cv2 4.13.0, is:
//   per threshold t: binary = gray > t; findContours(RETR_LIST, CHAIN_APPROX_NONE) = Suzuki-Abe
//   border following with a zero frame around the image; per contour: polygon moments, area /
//   inertia / convexity / colour filters, centre, median radius; then grouping across thresholds.
//
// GPU decomposition (all 17 thresholds of all frames of a chunk in flight at once). Border following is
// NOT replayed sequentially with pixel marks: blob_walk.cuh states (and tests/test_blob_walk_host.py
// checks against the Suzuki-Abe restatement) the equivalent mark-free formulation -- borders are the
// cycles of a successor function on (pixel, direction) states, and the raster scan discovers each cycle
// at its raster-first state -- so every border is followed by its own lane:
//   B1  blob_binarize_kernel   gray -> 17 bit planes per frame (one pass over the frame, 1 B/px read):
//                              a thread bit-slices 32 pixels once, each plane is then a handful of word ops
//   B2a blob_walk_kernel       persistent warps scan the planes tile by tile for candidate first pixels
//                              of components / holes (word ops); each lane takes a candidate and sends two
//                              walkers around its border in opposite directions, which either meet (the
//                              candidate is where OpenCV's scan discovers that border: a record with the
//                              border's length and exact area is emitted if the area passes the filter)
//                              or run into a state the scan meets earlier (not the start: dropped).
//                              Lanes refill from a per-warp list; no marks, no per-plane sequential pass.
//   B2b blob_points_kernel     one lane per kept border walks it once more: stores its points and sums the
//                              remaining Green's-theorem moments and the bounding box (exact integers).
//   B3  blob_contour_warp_kernel  one warp per kept border: convex-hull area from per-column extremes,
//                              centre, colour test, median point distance by radix selection on the IEEE
//                              bit patterns, all in the warp's slice of shared memory;
//       blob_contour_kernel    the same with one CTA and global scratch, for borders too long for that slice
//   host                       the remaining double arithmetic (inertia, convexity ratio) and the
//                              grouping across thresholds, a few hundred centres per frame, in the
//                              reference's operation order (no FMA contraction: see build.py), frames
//                              spread over host threads.
#include <cuda_runtime.h>
#include <float.h>
#include <limits.h>
#include <math.h>
#include <stdio.h>
#include <string.h>
#include <algorithm>
#include <thread>
#include <vector>

#include "kernels.cuh"
#include "blob_walk.cuh"

namespace mrgb200
{
namespace
{
using namespace blobwalk;

constexpr int kNThr = 17;                       // thresholds 50, 60, ..., 210 (min 50, max 220, step 10)
__host__ __device__ constexpr int thr_value(int k) { return 50 + 10 * k; }

struct BlobRecord
{
    int frame, thr, seq, n;                     // seq = position y*w + x at which the raster scan discovers the border
    unsigned pts_off;                           // first point, within the chunk's point region
    int sx, sy, sk;                             // the state the border is followed from (blob_walk.cuh)
    int xmin, xmax, ymin, ymax;
    long long a00, a10, a01, a20, a11, a02;     // Green's-theorem sums over the directed border edges
    long long hull2;                            // twice the convex hull's area
    double cx, cy, radius;
    int colour_ok, big;                         // big: left to the CTA-per-border kernel
};

struct BlobGeom
{
    int w, h, wpr;                              // pixels; storage words per bit-plane row (blobwalk::plane_wpr)
    int nframes;
    unsigned pts_cap;                           // points of the whole chunk
    unsigned rec_cap;
    unsigned queue_cap;                         // candidate starts of the whole chunk
    size_t plane_words;                         // storage words of one plane: (h + 2) rows of wpr words
    int origin;                                 // word offset of pixel (0,0) in a plane's storage
};
__device__ __forceinline__ const uint32_t* plane_ptr(const BlobGeom& g, const uint32_t* planes, int job)
{
    return planes + (size_t)job * g.plane_words + g.origin;
}

// device counters (uint32 each)
enum { kCntRecords = 0, kCntStatus = 1, kCntPoints = 2, kCntQueue = 3, kCntPointJob = 4, kCntQueueHead = 5, kCntWords = 8 };

// ------------------------------------------------------------------------------------------------
// B1: bit planes. plane(f,k)[y][wd] bit b = gray(f, y, 32*wd + b) > thr_k
// A thread owns 32 pixels: two 128-bit loads, bytes regrouped so that word j holds pixels j, j+8, j+16,
// j+24 (PRMT), then eight bit slices S_b (bit i = bit b of pixel i); "pixel > constant" on slices is one
// AND or OR per bit of the constant, 32 pixels at a time.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t slices_gt(const uint32_t (&S)[8], int t)
{
    uint32_t r = 0;
#pragma unroll
    for (int b = 0; b < 8; b++) r = ((t >> b) & 1) ? (S[b] & r) : (S[b] | r);
    return r;
}

__global__ void __launch_bounds__(128)
blob_binarize_kernel(FrameSet fs, BlobGeom g, uint32_t* __restrict__ planes, int aligned)
{
    // one thread per storage word: the frame of background words around the image is written here too
    const int ws = blockIdx.x * blockDim.x + threadIdx.x, ys = blockIdx.y, f = blockIdx.z;
    if (ws >= g.wpr) return;
    uint32_t* out = planes + (size_t)f * kNThr * g.plane_words + (size_t)ys * g.wpr + ws;
    const int wd = ws - 1, y = ys - 1;
    if (y < 0 || y >= g.h || wd < 0 || wd * 32 >= g.w)
    {
#pragma unroll
        for (int k = 0; k < kNThr; k++) out[(size_t)k * g.plane_words] = 0u;
        return;
    }
    const uint8_t* row = fs.base + (size_t)f * fs.frame_stride + (size_t)y * fs.pitch;
    const int x0 = wd * 32;
    uint32_t W[8];
    if (aligned)
    {
        uint4 a = make_uint4(0, 0, 0, 0), b = make_uint4(0, 0, 0, 0);
        if (x0 < g.w)      a = *reinterpret_cast<const uint4*>(row + x0);          // rows are multiples of 16 bytes long:
        if (x0 + 16 < g.w) b = *reinterpret_cast<const uint4*>(row + x0 + 16);     // a vector that starts inside ends inside
        W[0] = a.x; W[1] = a.y; W[2] = a.z; W[3] = a.w; W[4] = b.x; W[5] = b.y; W[6] = b.z; W[7] = b.w;
    }
    else
    {
#pragma unroll
        for (int j = 0; j < 8; j++)
        {
            uint32_t v = 0;
#pragma unroll
            for (int i = 0; i < 4; i++) if (x0 + 4*j + i < g.w) v |= (uint32_t)row[x0 + 4*j + i] << (8 * i);
            W[j] = v;
        }
    }
    if (x0 + 32 > g.w)
    {
        // pixels beyond the width are background in every plane
#pragma unroll
        for (int j = 0; j < 8; j++)
        {
            const int valid = g.w - (x0 + 4*j);
            W[j] &= valid >= 4 ? ~0u : (valid > 0 ? (1u << (8 * valid)) - 1u : 0u);
        }
    }
    uint32_t S[8];
#pragma unroll
    for (int b = 0; b < 8; b++) S[b] = 0;
#pragma unroll
    for (int j = 0; j < 8; j++)
    {
        // U = pixels j, j+8, j+16, j+24 in bytes 0..3
        const int a = j >> 2, bsel = j & 3;
        const uint32_t t1 = __byte_perm(W[a], W[a + 2], bsel | ((4 + bsel) << 4));
        const uint32_t t2 = __byte_perm(W[a + 4], W[a + 6], bsel | ((4 + bsel) << 4));
        const uint32_t U = __byte_perm(t1, t2, 0x5410);
#pragma unroll
        for (int b = 0; b < 8; b++)
        {
            const uint32_t sh = j >= b ? U << (j - b) : U >> (b - j);               // bit b of byte i -> bit 8i + j
            S[b] |= sh & (0x01010101u << j);
        }
    }
#pragma unroll
    for (int k = 0; k < kNThr; k++) out[(size_t)k * g.plane_words] = slices_gt(S, thr_value(k));
}

// ------------------------------------------------------------------------------------------------
// B2a: candidate starts (blob_scan_kernel) and their verification walks (blob_walk_kernel); see blob_walk.cuh
// ------------------------------------------------------------------------------------------------
constexpr int kTileRows  = 16;                  // a tile = 32 words (1024 pixels) x 16 rows of one plane, one warp
constexpr unsigned kFull = 0xffffffffu;
// queue entry: job << 32 | kind << 30 | y << 15 | x   (kind 0: first pixel of a component, 1: first pixel of a hole)

__global__ void __launch_bounds__(256)
blob_scan_kernel(BlobGeom g, const uint32_t* __restrict__ planes, unsigned long long* __restrict__ queue,
                 unsigned* __restrict__ counters, unsigned ntiles, int nstrips, int ncb)
{
    const int lane = threadIdx.x & 31;
    const unsigned nwarps = gridDim.x * (blockDim.x >> 5);
    const unsigned njobs = (unsigned)g.nframes * kNThr;
    // tile order: the first strip of every plane, then the second, ...: the longest borders (frame- and board-sized
    // outlines) start near the top of their planes, so they are at the head of the queue and are walked first
    for (unsigned t = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); t < ntiles; t += nwarps)
    {
        const int job = (int)(t % njobs);
        const int cb = (int)((t / njobs) % (unsigned)ncb), st = (int)(t / (njobs * (unsigned)ncb));
        const int y0 = st * kTileRows, rows = min(kTileRows, g.h - y0), wd = cb * 32 + lane;
        const uint32_t* B = plane_ptr(g, planes, job);
        const bool in = wd * 32 < g.w;                      // (words past the image are background; so are rows -1 and h)
        // the tile's rows -1 .. rows-1 of this lane's word column, all loads in flight at once; the words of the
        // neighbouring columns come from the neighbouring lanes (lanes 0 and 31 load theirs)
        uint32_t c[kTileRows + 1], el[kTileRows + 1], er[kTileRows + 1];
#pragma unroll
        for (int r = 0; r <= kTileRows; r++) c[r] = (in && r <= rows) ? B[(y0 + r - 1) * g.wpr + wd] : 0u;
        if (lane == 0 || lane == 31)
        {
            const int wn = lane == 0 ? wd - 1 : wd + 1;
#pragma unroll
            for (int r = 0; r <= kTileRows; r++) el[r] = (in && r <= rows) ? B[(y0 + r - 1) * g.wpr + wn] : 0u;
        }
#pragma unroll
        for (int r = 0; r <= kTileRows; r++)
        {
            const uint32_t l = __shfl_up_sync(kFull, c[r], 1), rr = __shfl_down_sync(kFull, c[r], 1);
            er[r] = lane == 31 ? el[r] : rr;
            el[r] = lane == 0 ? el[r] : l;
        }
        uint32_t outer[kTileRows], hole[kTileRows];
        unsigned cnt = 0;
#pragma unroll
        for (int r = 0; r < kTileRows; r++)
        {
            candidate_masks(c[r + 1], el[r + 1], c[r], el[r], er[r], &outer[r], &hole[r]);
            if (r >= rows) { outer[r] = 0; hole[r] = 0; }
            cnt += __popc(outer[r]) + __popc(hole[r]);
        }
        if (__ballot_sync(kFull, cnt != 0))
        {
            unsigned incl = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const unsigned v = __shfl_up_sync(kFull, incl, o); if (lane >= o) incl += v; }
            unsigned base = 0;
            if (lane == 31) base = atomicAdd(&counters[kCntQueue], incl);
            base = __shfl_sync(kFull, base, 31);
            unsigned at = base + incl - cnt;
#pragma unroll
            for (int r = 0; r < kTileRows; r++)
            {
                if ((outer[r] | hole[r]) == 0) continue;
                const unsigned long long hi = (unsigned long long)job << 32 | (unsigned)(y0 + r) << 15 | (unsigned)(wd * 32);
                uint32_t m = outer[r];
                while (m) { const int bb = __ffs(m) - 1; m &= m - 1; if (at < g.queue_cap) queue[at] = hi | (unsigned)bb; at++; }
                m = hole[r];
                while (m) { const int bb = __ffs(m) - 1; m &= m - 1; if (at < g.queue_cap) queue[at] = hi | 1ull << 30 | (unsigned)bb; at++; }
            }
        }
    }
}

__global__ void __launch_bounds__(128)
blob_walk_kernel(BlobGeom g, const uint32_t* __restrict__ planes, const unsigned long long* __restrict__ queue,
                 BlobRecord* __restrict__ recs, unsigned* __restrict__ counters)
{
    const int lane = threadIdx.x & 31;
    const unsigned qn_all = counters[kCntQueue];
    if (qn_all > g.queue_cap) { if (threadIdx.x == 0 && blockIdx.x == 0) atomicExch(&counters[kCntStatus], 1u); return; }
    const unsigned qn = qn_all;
    const long long max_steps = 4LL * g.w * g.h + 16;

    bool active = false, more = true;
    PlaneRef P; P.B = planes; P.w = g.w; P.h = g.h; P.wpr = g.wpr;
    BitWindow F, Bk;
    int fx = 0, fy = 0, fk = 0, bx = 0, by = 0, bk = 0, pos = 0, cnt = 0, sx = 0, sy = 0, sk = 0, job = 0;
    long long a00 = 0;

    for (;;)
    {
        // idle lanes take the queue's next candidates
        const unsigned idle = __ballot_sync(kFull, !active);
        if (idle && more)
        {
            unsigned base = 0;
            if (lane == 0) base = atomicAdd(&counters[kCntQueueHead], (unsigned)__popc(idle));
            base = __shfl_sync(kFull, base, 0);
            more = base < qn;
            const unsigned my = base + __popc(idle & ((1u << lane) - 1u));
            if (!active && my < qn)
            {
                const unsigned long long e = queue[my];
                const int x = (int)(e & 0x7FFFu), y = (int)((e >> 15) & 0x7FFFu);
                job = (int)(e >> 32); P.B = plane_ptr(g, planes, job);
                F.init(); Bk.init();
                pos = y * g.w + x;
                bool ok = true;
                if ((e >> 30) & 1u) { sx = x - 1; sy = y; sk = 1; }
                else { sx = x; sy = y; ok = outer_start(P, F, x, y, &sk); }          // false: isolated pixel, area 0
                if (ok) { fx = bx = sx; fy = by = sy; fk = bk = sk; a00 = 0; cnt = 0; active = true; }
            }
        }
        if (!__any_sync(kFull, active)) { if (!more) break; continue; }

#pragma unroll 1
        for (int it = 0; it < 8; it++)
        {
            if (active)
            {
                // Both walkers step every time, as two independent dependency chains; the backward step is thrown
                // away when the forward one already decided (order of the tests as in blobwalk::verify_start).
                int discf, discb;
                const int px = fx, py = fy, qx = bx, qy = by, qk = bk;
                step_fwd(P, F, fx, fy, fk, &discf);
                step_bwd(P, Bk, bx, by, bk, &discb);
                a00 += (long long)(px * fy - fx * py); cnt++;
                bool drop = discf >= 0 && discf < pos;
                bool met = !drop && fx == qx && fy == qy && fk == qk;
                if (!drop && !met)
                {
                    a00 += (long long)(bx * qy - qx * by); cnt++;
                    drop = discb >= 0 && discb < pos;
                    met = !drop && fx == bx && fy == by && fk == bk;
                }
                if (!drop && !met && cnt > max_steps) { atomicExch(&counters[kCntStatus], 2u); drop = true; }   // cannot happen
                if (drop) active = false;
                else if (met)
                {
                    active = false;
                    // filterByArea: m00 = |a00| / 2 in [20, 80000) -- exact in integers. Everything else is dropped here.
                    const long long aa = a00 < 0 ? -a00 : a00;
                    if (aa >= 40 && aa < 160000)
                    {
                        const unsigned idx = atomicAdd(&counters[kCntRecords], 1u);
                        const unsigned off = atomicAdd(&counters[kCntPoints], (unsigned)cnt);
                        if (idx >= g.rec_cap || off > g.pts_cap || (unsigned)cnt > g.pts_cap - off) atomicExch(&counters[kCntStatus], 1u);
                        else
                        {
                            BlobRecord r;
                            r.frame = job / kNThr; r.thr = job % kNThr; r.seq = pos; r.n = cnt; r.pts_off = off;
                            r.sx = sx; r.sy = sy; r.sk = sk;
                            r.xmin = r.xmax = r.ymin = r.ymax = 0;
                            r.a00 = a00; r.a10 = r.a01 = r.a20 = r.a11 = r.a02 = 0;
                            r.hull2 = 0; r.cx = r.cy = r.radius = 0; r.colour_ok = 0; r.big = 0;
                            recs[idx] = r;
                        }
                    }
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// B2b: points, remaining moments and bounding box of the kept borders, one lane per border
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
blob_points_kernel(BlobGeom g, const uint32_t* __restrict__ planes, BlobRecord* __restrict__ recs,
                   unsigned* __restrict__ counters, uint32_t* __restrict__ pts)
{
    const int lane = threadIdx.x & 31;
    const unsigned nrec = min(counters[kCntRecords], g.rec_cap);
    if (counters[kCntStatus]) return;
    PlaneRef P; P.B = planes; P.w = g.w; P.h = g.h; P.wpr = g.wpr;
    BitWindow F;
    bool active = false, more = true;
    unsigned ri = 0; int x = 0, y = 0, k = 0, i = 0, n = 0;
    uint32_t* out = pts;
    long long t10 = 0, t01 = 0, t20 = 0, t11 = 0, t02 = 0;
    int bx0 = 0, bx1 = 0, by0 = 0, by1 = 0;
    for (;;)
    {
        const unsigned idle = __ballot_sync(kFull, !active);
        if (idle && more)
        {
            unsigned base = 0;
            if (lane == 0) base = atomicAdd(&counters[kCntPointJob], (unsigned)__popc(idle));
            base = __shfl_sync(kFull, base, 0);
            more = base < nrec;
            if (!active)
            {
                ri = base + __popc(idle & ((1u << lane) - 1u));
                if (ri < nrec)
                {
                    const BlobRecord& r = recs[ri];
                    P.B = plane_ptr(g, planes, r.frame * kNThr + r.thr);
                    F.init();
                    x = r.sx; y = r.sy; k = r.sk; n = r.n; i = 0; out = pts + r.pts_off;
                    t10 = t01 = t20 = t11 = t02 = 0;
                    bx0 = INT_MAX; bx1 = INT_MIN; by0 = INT_MAX; by1 = INT_MIN;
                    active = true;
                }
            }
        }
        if (!__any_sync(kFull, active)) { if (!more) break; continue; }
#pragma unroll 1
        for (int it = 0; it < 32; it++)
        {
            if (active)
            {
                out[i] = (uint32_t)x | ((uint32_t)y << 16);
                bx0 = min(bx0, x); bx1 = max(bx1, x); by0 = min(by0, y); by1 = max(by1, y);
                const long long xp = x, yp = y;
                int disc;
                step_fwd(P, F, x, y, k, &disc);
                // edge (xp,yp) -> (x,y): the terms of cv::moments' contour sums, exact integers
                const long long xc = x, yc = y;
                const long long dxy = xp * yc - xc * yp, xs = xp + xc, ys2 = yp + yc;
                t10 += dxy * xs; t01 += dxy * ys2;
                t20 += dxy * (xp * xs + xc * xc);
                t11 += dxy * (xp * (ys2 + yp) + xc * (ys2 + yc));
                t02 += dxy * (yp * ys2 + yc * yc);
                if (++i == n)
                {
                    BlobRecord& r = recs[ri];
                    r.a10 = t10; r.a01 = t01; r.a20 = t20; r.a11 = t11; r.a02 = t02;
                    r.xmin = bx0; r.xmax = bx1; r.ymin = by0; r.ymax = by1;
                    active = false;
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// B3: hull area, colour test, median radius of one kept border
// ------------------------------------------------------------------------------------------------
// convex hull of the border = hull of the per-column extremes (every column of the box holds a border
// point). Monotone chains over the columns; the area comes out as an exact integer: the polygon is the lower
// chain left -> right, up the last column, the upper chain right -> left, down the first column.
// One thread per chain (stk: 2 * bw ints each).
__device__ long long hull_chain_sum(const int* ys, bool lower, int* stk, int bw)
{
    int k = 0;
    for (int i = 0; i < bw; i++)
    {
        const long long x = i, y = ys[i];
        while (k >= 2)
        {
            const long long ox = stk[2*(k-2)], oy = stk[2*(k-2)+1], ax = stk[2*(k-1)], ay = stk[2*(k-1)+1];
            const long long cr = (ax - ox) * (y - oy) - (ay - oy) * (x - ox);
            if (lower ? cr <= 0 : cr >= 0) k--; else break;
        }
        stk[2*k] = i; stk[2*k+1] = (int)y; k++;
    }
    long long s = 0;
    for (int i = 0; i + 1 < k; i++)
        s += (long long)stk[2*i] * stk[2*i+3] - (long long)stk[2*i+2] * stk[2*i+1];
    return s;
}
__device__ __forceinline__ long long hull_area2_from_chains(long long sl, long long su, const int* ylo, const int* yhi, int bw)
{
    const long long xe = bw - 1;
    const long long a2 = sl + xe * ((long long)yhi[bw-1] - ylo[bw-1]) - su;      // first column: x = 0 contributes nothing
    return a2 < 0 ? -a2 : a2;
}

// centre with the rounding of cv::moments / SimpleBlobDetector (m = a * (+-1/2, +-1/6), c = m10 / m00) and the colour
// filter (the binary image must be 0 at (cvRound(cy), cvRound(cx))). One thread. The host rejects a record whose
// colour test fails whatever its hull is, so the hull (and the radius) are only worked out for the others.
__device__ bool centre_and_colour(BlobRecord& r, const BlobGeom& g, const uint32_t* __restrict__ planes)
{
    const double sgn = r.a00 > 0 ? 1.0 : -1.0;
    const double m00 = __dmul_rn((double)r.a00, sgn * 0.5);
    const double m10 = __dmul_rn((double)r.a10, sgn * 0.16666666666666666666666666666667);
    const double m01 = __dmul_rn((double)r.a01, sgn * 0.16666666666666666666666666666667);
    const double cx = __ddiv_rn(m10, m00), cy = __ddiv_rn(m01, m00);
    r.cx = cx; r.cy = cy;
    const int rx = __double2int_rn(cx), ry = __double2int_rn(cy);
    int ok = 0;
    if (rx >= 0 && rx < g.w && ry >= 0 && ry < g.h)
    {
        const uint32_t* B = plane_ptr(g, planes, r.frame * kNThr + r.thr);
        ok = !((B[ry * g.wpr + (rx >> 5)] >> (rx & 31)) & 1u);
    }
    r.colour_ok = ok;
    return ok != 0;
}
// The host also rejects the record if area / hullArea < 0.95f; where that is certain (a ratio below 0.94: far from
// any rounding question) the median radius is not needed.
__device__ __forceinline__ bool radius_needed(const BlobRecord& r)
{
    const long long aa = r.a00 < 0 ? -r.a00 : r.a00;
    return aa * 100 >= r.hull2 * 94;
}

__device__ __forceinline__ unsigned long long dist2_key(double cx, double cy, uint32_t q)
{
    const double dx = __dsub_rn(cx, (double)(int)(q & 0xFFFF)), dy = __dsub_rn(cy, (double)(int)(q >> 16));
    return (unsigned long long)__double_as_longlong(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
}

constexpr int kW3Keys = 1536, kW3Cols = 512, kW3Warps = 6;
struct WarpScratch
{
    unsigned long long keys[kW3Keys];           // the hull's two chain stacks (2 * kW3Cols ints each) live here before the keys do
    int ylo[kW3Cols], yhi[kW3Cols];
    unsigned hist[256];
};

// One warp per border whose points (<= kW3Keys) and columns (<= kW3Cols) fit the warp's slice of shared memory.
__global__ void __launch_bounds__(kW3Warps * 32)
blob_contour_warp_kernel(BlobGeom g, const uint32_t* __restrict__ planes, const uint32_t* __restrict__ pts,
                         BlobRecord* recs, const unsigned* __restrict__ counters)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    WarpScratch& S = reinterpret_cast<WarpScratch*>(smem_raw)[threadIdx.x >> 5];
    const int lane = threadIdx.x & 31;
    if (counters[kCntStatus]) return;
    const unsigned nrec = min(counters[kCntRecords], g.rec_cap);
    const unsigned nwarps = gridDim.x * kW3Warps;
    for (unsigned ri = blockIdx.x * kW3Warps + (threadIdx.x >> 5); ri < nrec; ri += nwarps)
    {
        BlobRecord& r = recs[ri];
        const int n = r.n, bw = r.xmax - r.xmin + 1;
        if (n > kW3Keys || bw > kW3Cols) { if (lane == 0) r.big = 1; continue; }
        const uint32_t* p = pts + r.pts_off;
        int need = 0;
        if (lane == 0) need = centre_and_colour(r, g, planes);
        need = __shfl_sync(kFull, need, 0);
        if (!need) continue;
        for (int i = lane; i < bw; i += 32) { S.ylo[i] = INT_MAX; S.yhi[i] = INT_MIN; }
        __syncwarp();
        for (int i = lane; i < n; i += 32)
        {
            const uint32_t q = p[i];
            const int x = (int)(q & 0xFFFF) - r.xmin, y = (int)(q >> 16);
            atomicMin(&S.ylo[x], y); atomicMax(&S.yhi[x], y);
        }
        __syncwarp();
        // lower chain on lane 0, upper chain on lane 1
        long long cs = 0;
        if (lane < 2) cs = hull_chain_sum(lane ? S.yhi : S.ylo, lane == 0, reinterpret_cast<int*>(S.keys) + lane * 2 * kW3Cols, bw);
        const long long su = __shfl_sync(kFull, cs, 1);
        if (lane == 0)
        {
            r.hull2 = hull_area2_from_chains(cs, su, S.ylo, S.yhi, bw);
            need = radius_needed(r);
        }
        need = __shfl_sync(kFull, need, 0);
        if (!need) continue;
        const double cx = __shfl_sync(kFull, r.cx, 0), cy = __shfl_sync(kFull, r.cy, 0);   // (lane 0 wrote them; same value for all)
        __syncwarp();
        for (int i = lane; i < n; i += 32) S.keys[i] = dist2_key(cx, cy, p[i]);
        __syncwarp();
        // median of the point distances to the centre: order statistics (n-1)/2 and n/2 of d2 = dx*dx + dy*dy,
        // selected on the bit patterns (non-negative doubles order like integers), one byte per pass
        double dsel[2];
        for (int which = 0; which < 2; which++)
        {
            unsigned rank = which == 0 ? (unsigned)(n - 1) / 2 : (unsigned)n / 2;
            if (which == 1 && rank == (unsigned)(n - 1) / 2) { dsel[1] = dsel[0]; break; }
            unsigned long long prefix = 0;
            for (int pass = 7; pass >= 0; pass--)
            {
                for (int i = lane; i < 256; i += 32) S.hist[i] = 0;
                __syncwarp();
                for (int i = lane; i < n; i += 32)
                {
                    const unsigned long long key = S.keys[i];
                    if (pass == 7 || (key >> (8 * (pass + 1))) == (prefix >> (8 * (pass + 1))))
                        atomicAdd(&S.hist[(key >> (8 * pass)) & 255], 1u);
                }
                __syncwarp();
                // lane l owns bins 8l .. 8l+7
                unsigned h[8], mine = 0;
#pragma unroll
                for (int j = 0; j < 8; j++) { h[j] = S.hist[8 * lane + j]; mine += h[j]; }
                unsigned incl = mine;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const unsigned v = __shfl_up_sync(kFull, incl, o); if (lane >= o) incl += v; }
                const unsigned excl = incl - mine;
                const bool here = rank >= excl && rank < incl;
                unsigned b = 0, rk = 0;
                if (here)
                {
                    rk = rank - excl;
                    bool found = false;
#pragma unroll
                    for (int j = 0; j < 8; j++)
                        if (!found) { if (rk < h[j]) { found = true; b = 8 * lane + j; } else rk -= h[j]; }
                }
                const int src = __ffs(__ballot_sync(kFull, here)) - 1;
                b = __shfl_sync(kFull, b, src); rank = __shfl_sync(kFull, rk, src);
                prefix |= (unsigned long long)b << (8 * pass);
                __syncwarp();
            }
            dsel[which] = __dsqrt_rn(__longlong_as_double((long long)prefix));
        }
        if (lane == 0) r.radius = __ddiv_rn(__dadd_rn(dsel[0], dsel[1]), 2.0);
        __syncwarp();
    }
}

// The same for the borders the warp kernel left (big = 1): one CTA each, per-column extremes and the chain stack
// in global scratch, keys recomputed in every pass.
constexpr int kB3Threads = 128;

__global__ void __launch_bounds__(kB3Threads)
blob_contour_kernel(BlobGeom g, const uint32_t* __restrict__ planes, const uint32_t* __restrict__ pts,
                    BlobRecord* recs, const unsigned* __restrict__ counters, int* scratch, int scratch_stride)
{
    __shared__ unsigned hist[256];
    __shared__ unsigned long long sel_prefix;
    __shared__ unsigned sel_rank;
    __shared__ int s_need_radius;
    if (counters[kCntStatus]) return;
    const unsigned nrec = min(counters[kCntRecords], g.rec_cap);
    int* ylo = scratch + (size_t)blockIdx.x * scratch_stride;       // per column of the bounding box
    int* yhi = ylo + g.w;
    int* stk = yhi + g.w;                                           // hull chain: (x, y) pairs
    for (unsigned ri = blockIdx.x; ri < nrec; ri += gridDim.x)
    {
        BlobRecord& r = recs[ri];
        if (!r.big) continue;
        const uint32_t* p = pts + r.pts_off;
        const int n = r.n;
        const int bw = r.xmax - r.xmin + 1;
        if (threadIdx.x == 0) s_need_radius = centre_and_colour(r, g, planes);
        __syncthreads();
        if (!s_need_radius) { __syncthreads(); continue; }
        for (int i = threadIdx.x; i < bw; i += kB3Threads) { ylo[i] = INT_MAX; yhi[i] = INT_MIN; }
        __syncthreads();
        for (int i = threadIdx.x; i < n; i += kB3Threads)
        {
            const uint32_t q = p[i];
            const int x = (int)(q & 0xFFFF) - r.xmin, y = (int)(q >> 16);
            atomicMin(&ylo[x], y); atomicMax(&yhi[x], y);
        }
        __syncthreads();
        if (threadIdx.x == 0)
        {
            const long long sl = hull_chain_sum(ylo, true, stk, bw), su = hull_chain_sum(yhi, false, stk, bw);
            r.hull2 = hull_area2_from_chains(sl, su, ylo, yhi, bw);
            s_need_radius = radius_needed(r);
        }
        __syncthreads();
        if (!s_need_radius) { __syncthreads(); continue; }
        const double cx = r.cx, cy = r.cy;
        double dsel[2];
        for (int which = 0; which < 2; which++)
        {
            const unsigned rank0 = which == 0 ? (unsigned)(n - 1) / 2 : (unsigned)n / 2;
            if (which == 1 && rank0 == (unsigned)(n - 1) / 2) { dsel[1] = dsel[0]; break; }
            if (threadIdx.x == 0) { sel_prefix = 0; sel_rank = rank0; }
            for (int pass = 7; pass >= 0; pass--)
            {
                for (int i = threadIdx.x; i < 256; i += kB3Threads) hist[i] = 0;
                __syncthreads();
                const unsigned long long prefix = sel_prefix;
                for (int i = threadIdx.x; i < n; i += kB3Threads)
                {
                    const unsigned long long key = dist2_key(cx, cy, p[i]);
                    if (pass == 7 || (key >> (8 * (pass + 1))) == (prefix >> (8 * (pass + 1))))
                        atomicAdd(&hist[(key >> (8 * pass)) & 255], 1u);
                }
                __syncthreads();
                if (threadIdx.x == 0)
                {
                    unsigned rk = sel_rank, b = 0;
                    while (rk >= hist[b]) { rk -= hist[b]; b++; }
                    sel_rank = rk;
                    sel_prefix = prefix | ((unsigned long long)b << (8 * pass));
                }
                __syncthreads();
            }
            dsel[which] = __dsqrt_rn(__longlong_as_double((long long)sel_prefix));
            __syncthreads();
        }
        if (threadIdx.x == 0) r.radius = __ddiv_rn(__dadd_rn(dsel[0], dsel[1]), 2.0);
        __syncthreads();
    }
}

struct Center { double x, y, radius, confidence; };

// SimpleBlobDetector::findBlobs' remaining filters on one border (inertia, convexity, colour)
bool center_from_record(const BlobRecord& r, Center* c)
{
    const double a00 = (double)r.a00, a10 = (double)r.a10, a01 = (double)r.a01, a20 = (double)r.a20, a11 = (double)r.a11, a02 = (double)r.a02;
    double db1_2, db1_6, db1_12, db1_24;
    if (a00 > 0) { db1_2 = 0.5; db1_6 = 0.16666666666666666666666666666667; db1_12 = 0.083333333333333333333333333333333; db1_24 = 0.041666666666666666666666666666667; }
    else { db1_2 = -0.5; db1_6 = -0.16666666666666666666666666666667; db1_12 = -0.083333333333333333333333333333333; db1_24 = -0.041666666666666666666666666666667; }
    const double m00 = a00 * db1_2, m10 = a10 * db1_6, m01 = a01 * db1_6, m20 = a20 * db1_12, m11 = a11 * db1_24, m02 = a02 * db1_12;
    const double inv_m00 = 1. / m00, cx = m10 * inv_m00, cy = m01 * inv_m00;
    const double mu20 = m20 - m10 * cx, mu11 = m11 - m10 * cy, mu02 = m02 - m01 * cy;
    c->confidence = 1;
    if (m00 < 20.0f || m00 >= 80000.0f) return false;
    {
        const double t = 2 * mu11, d = mu20 - mu02;
        const double denominator = sqrt(t * t + d * d);
        double ratio;
        if (denominator > 1e-2)
        {
            const double cosmin = (mu20 - mu02) / denominator, sinmin = 2 * mu11 / denominator;
            const double cosmax = -cosmin, sinmax = -sinmin;
            const double imin = 0.5 * (mu20 + mu02) - 0.5 * (mu20 - mu02) * cosmin - mu11 * sinmin;
            const double imax = 0.5 * (mu20 + mu02) - 0.5 * (mu20 - mu02) * cosmax - mu11 * sinmax;
            ratio = imin / imax;
        }
        else ratio = 1;
        if (ratio < 0.1f || ratio >= FLT_MAX) return false;
        c->confidence = ratio * ratio;
    }
    {
        const double hullArea = 0.5 * (double)r.hull2;
        if (fabs(hullArea) < DBL_EPSILON) return false;
        const double ratio = m00 / hullArea;
        if (ratio < 0.95f || ratio >= FLT_MAX) return false;
    }
    if (!r.colour_ok) return false;
    c->x = r.cx; c->y = r.cy; c->radius = r.radius;
    return true;
}

// SimpleBlobDetector::detect's grouping across thresholds + find_blobs.cc:40-41, for one frame.
// order: this frame's records sorted by (thr ascending, seq DESCENDING): OpenCV hands contours over
// in reverse discovery order.
int group_frame(const BlobRecord* all, const unsigned* order, int nrec, int32_t* xy_out, int max_points)
{
    std::vector<std::vector<Center>> centers;
    int i = 0;
    while (i < nrec)
    {
        const int thr = all[order[i]].thr;
        std::vector<std::vector<Center>> fresh;
        for (; i < nrec && all[order[i]].thr == thr; i++)
        {
            Center c;
            if (!center_from_record(all[order[i]], &c)) continue;
            bool is_new = true;
            for (size_t j = 0; j < centers.size(); j++)
            {
                const Center& mid = centers[j][centers[j].size() / 2];
                const double dx = mid.x - c.x, dy = mid.y - c.y, dist = sqrt(dx * dx + dy * dy);
                is_new = dist >= 5.0f && dist >= mid.radius && dist >= c.radius;
                if (!is_new)
                {
                    centers[j].push_back(c);
                    size_t k = centers[j].size() - 1;
                    while (k > 0 && c.radius < centers[j][k - 1].radius) { centers[j][k] = centers[j][k - 1]; k--; }
                    centers[j][k] = c;
                    break;
                }
            }
            if (is_new) fresh.push_back(std::vector<Center>(1, c));
        }
        for (auto& v : fresh) centers.push_back(v);
    }
    int n = 0;
    for (auto& grp : centers)
    {
        if (grp.size() < 2) continue;
        double sx = 0, sy = 0, normalizer = 0;
        for (auto& c : grp) { sx += c.confidence * c.x; sy += c.confidence * c.y; normalizer += c.confidence; }
        const double inv = 1. / normalizer;
        sx *= inv; sy *= inv;
        const float fx = (float)sx, fy = (float)sy;
        if (n < max_points)
        {
            const float px = fx * 1000, py = fy * 1000;      // float * int -> float; the + 0.5 below is double
            xy_out[2*n] = (int)(px + 0.5); xy_out[2*n + 1] = (int)(py + 0.5);
        }
        n++;
    }
    return n;
}
}   // namespace

struct BlobWorkspace
{
    void* planes = nullptr; void* pts = nullptr; void* recs = nullptr;
    void* scratch = nullptr; void* counters = nullptr; void* queue = nullptr;
    size_t planes_b = 0, pts_b = 0, recs_b = 0, scratch_b = 0, queue_b = 0;
    unsigned pts_per_frame = 1u << 21, rec_per_job = 512, queue_per_frame = 1u << 18;
    BlobRecord* host_recs = nullptr; size_t host_recs_cap = 0;      // pinned
    unsigned* host_counters = nullptr;                              // pinned
    std::vector<unsigned> order, first;
    bool warp_smem_set = false;
    // the chunk in flight (blob_enqueue .. blob_finish)
    bool pending = false;
    BlobGeom g; cudaStream_t stream = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr, done = nullptr;
};

BlobWorkspace* blob_workspace_create() { return new BlobWorkspace(); }
void blob_workspace_destroy(BlobWorkspace* ws)
{
    if (!ws) return;
    cudaFree(ws->planes); cudaFree(ws->pts); cudaFree(ws->recs);
    cudaFree(ws->scratch); cudaFree(ws->counters); cudaFree(ws->queue);
    if (ws->host_recs) cudaFreeHost(ws->host_recs);
    if (ws->host_counters) cudaFreeHost(ws->host_counters);
    if (ws->e0) cudaEventDestroy(ws->e0);
    if (ws->e1) cudaEventDestroy(ws->e1);
    if (ws->done) cudaEventDestroy(ws->done);
    delete ws;
}

static cudaError_t grow(void** p, size_t* have, size_t want)
{
    if (want <= *have) return cudaSuccess;
    if (*p) cudaFree(*p);
    *p = nullptr; *have = 0;
    cudaError_t e = cudaMalloc(p, want);
    if (e == cudaSuccess) *have = want;
    return e;
}

#define BLOB_TRY(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { \
    fprintf(stderr, "%s:%d in %s(): CUDA failure '%s' in " #expr ". Sorry.\n", __FILE__, __LINE__, __func__, cudaGetErrorString(_e)); return -1; } } while (0)

void blob_workspace_reset_capacity(BlobWorkspace* ws) { ws->pts_per_frame = 1u << 21; ws->rec_per_job = 512; ws->queue_per_frame = 1u << 18; }

// Enqueues the kernels of one chunk of device-resident frames on `stream` (nothing is waited for; the frames may
// be overwritten once the first kernel has run, i.e. after whatever is enqueued next on `stream`). Returns 0 / -1.
int blob_enqueue(BlobWorkspace* ws, const FrameSet& fs, cudaStream_t stream)
{
    if (ws->pending) { fprintf(stderr, "%s:%d in %s(): a chunk is already in flight on this workspace. Sorry.\n", __FILE__, __LINE__, __func__); return -1; }
    const int n = fs.nframes;
    BlobGeom& g = ws->g;
    g.w = fs.w; g.h = fs.h; g.wpr = plane_wpr(fs.w); g.nframes = n;
    g.plane_words = (size_t)plane_rows(fs.h) * g.wpr; g.origin = plane_origin(fs.w);
    ws->stream = stream;
    if (n <= 0 || fs.w <= 0 || fs.h <= 0) { g.nframes = 0; ws->pending = true; return 0; }
    BLOB_TRY(grow(&ws->planes, &ws->planes_b, (size_t)n * kNThr * g.plane_words * 4));
    if (!ws->counters) BLOB_TRY(cudaMalloc(&ws->counters, kCntWords * 4));
    if (!ws->host_counters) BLOB_TRY(cudaMallocHost((void**)&ws->host_counters, kCntWords * 4));
    if (!ws->e0)
    {
        BLOB_TRY(cudaEventCreate(&ws->e0)); BLOB_TRY(cudaEventCreate(&ws->e1));
        BLOB_TRY(cudaEventCreateWithFlags(&ws->done, cudaEventDisableTiming));
    }
    const int b3_blocks = 148 * 4;
    const int scratch_stride = 4 * fs.w + 8;
    BLOB_TRY(grow(&ws->scratch, &ws->scratch_b, (size_t)b3_blocks * scratch_stride * sizeof(int)));
    if (!ws->warp_smem_set)
    {
        BLOB_TRY(cudaFuncSetAttribute(blob_contour_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kW3Warps * sizeof(WarpScratch))));
        ws->warp_smem_set = true;
    }
    const int aligned = ((uintptr_t)fs.base % 16 == 0) && (fs.pitch % 16 == 0) && (fs.frame_stride % 16 == 0 || n == 1);
    // tiles of the candidate scan: 32 words x kTileRows rows
    const int ncb = ((g.w + 31) / 32 + 31) / 32, nstrips = (g.h + kTileRows - 1) / kTileRows;
    const unsigned long long ntiles64 = (unsigned long long)n * kNThr * nstrips * ncb;
    if (ntiles64 >= 0xFFFFFFFFull) { fprintf(stderr, "%s:%d in %s(): too many frames in one blob chunk. Sorry.\n", __FILE__, __LINE__, __func__); return -1; }

    const unsigned long long want_pts = (unsigned long long)ws->pts_per_frame * n;
    g.pts_cap = (unsigned)std::min<unsigned long long>(want_pts, 0xFFFFFFF0ull);
    g.rec_cap = (unsigned)std::min<unsigned long long>((unsigned long long)ws->rec_per_job * n * kNThr, 0x7FFFFFFFull);
    g.queue_cap = (unsigned)std::min<unsigned long long>((unsigned long long)ws->queue_per_frame * n, 0x7FFFFFFFull);
    BLOB_TRY(grow(&ws->queue, &ws->queue_b, (size_t)g.queue_cap * 8));
    BLOB_TRY(grow(&ws->pts, &ws->pts_b, (size_t)g.pts_cap * 4));
    BLOB_TRY(grow(&ws->recs, &ws->recs_b, (size_t)g.rec_cap * sizeof(BlobRecord)));
    unsigned* counters = (unsigned*)ws->counters;
    BLOB_TRY(cudaMemsetAsync(ws->counters, 0, kCntWords * 4, stream));
    BLOB_TRY(cudaEventRecord(ws->e0, stream));
    blob_binarize_kernel<<<dim3((g.wpr + 127) / 128, plane_rows(g.h), n), 128, 0, stream>>>(fs, g, (uint32_t*)ws->planes, aligned);
    blob_scan_kernel<<<148 * 8, 256, 0, stream>>>(g, (const uint32_t*)ws->planes, (unsigned long long*)ws->queue, counters,
                                                  (unsigned)ntiles64, nstrips, ncb);
    blob_walk_kernel<<<148 * 6, 128, 0, stream>>>(g, (const uint32_t*)ws->planes, (const unsigned long long*)ws->queue,
                                                  (BlobRecord*)ws->recs, counters);
    blob_points_kernel<<<148 * 8, 128, 0, stream>>>(g, (const uint32_t*)ws->planes, (BlobRecord*)ws->recs, counters, (uint32_t*)ws->pts);
    blob_contour_warp_kernel<<<148 * 2, kW3Warps * 32, kW3Warps * sizeof(WarpScratch), stream>>>(
        g, (const uint32_t*)ws->planes, (const uint32_t*)ws->pts, (BlobRecord*)ws->recs, counters);
    blob_contour_kernel<<<b3_blocks, kB3Threads, 0, stream>>>(g, (const uint32_t*)ws->planes, (const uint32_t*)ws->pts,
                                                              (BlobRecord*)ws->recs, counters, (int*)ws->scratch, scratch_stride);
    BLOB_TRY(cudaEventRecord(ws->e1, stream));
    BLOB_TRY(cudaGetLastError());
    BLOB_TRY(cudaMemcpyAsync(ws->host_counters, ws->counters, kCntWords * 4, cudaMemcpyDeviceToHost, stream));
    BLOB_TRY(cudaEventRecord(ws->done, stream));
    ws->pending = true;
    return 0;
}

// Waits for the chunk enqueued last, then groups its records on the host. xy_out: HOST [nframes][max_points][2]
// int32 (scaled by 1000), counts_out: HOST [nframes]. ms_out (optional): device time of the kernels is ADDED.
// Returns 0, -1 on a CUDA failure, or 1 when the chunk's scratch overflowed (nothing was produced).
int blob_finish(BlobWorkspace* ws, int32_t* xy_out, int32_t* counts_out, int max_points, float* ms_out)
{
    if (!ws->pending) return -1;
    ws->pending = false;
    const BlobGeom& g = ws->g;
    const int n = g.nframes;
    for (int i = 0; i < n; i++) counts_out[i] = 0;
    if (n <= 0) return 0;
    BLOB_TRY(cudaEventSynchronize(ws->done));
    if (ms_out) { float t = 0; cudaEventElapsedTime(&t, ws->e0, ws->e1); *ms_out += t; }
    const unsigned* hc = ws->host_counters;
    if (hc[kCntStatus] == 2) { fprintf(stderr, "%s:%d in %s(): a border walk did not close. Sorry.\n", __FILE__, __LINE__, __func__); return -1; }
    if (hc[kCntStatus] != 0 || hc[kCntRecords] > g.rec_cap) return 1;
    const unsigned nrec = hc[kCntRecords];
    if (!nrec) return 0;

    if (ws->host_recs_cap < nrec)
    {
        if (ws->host_recs) cudaFreeHost(ws->host_recs);
        ws->host_recs = nullptr; ws->host_recs_cap = 0;
        const size_t cap = (size_t)nrec + nrec / 2 + 1024;
        BLOB_TRY(cudaMallocHost((void**)&ws->host_recs, cap * sizeof(BlobRecord)));
        ws->host_recs_cap = cap;
    }
    BLOB_TRY(cudaMemcpyAsync(ws->host_recs, ws->recs, sizeof(BlobRecord) * nrec, cudaMemcpyDeviceToHost, ws->stream));
    BLOB_TRY(cudaStreamSynchronize(ws->stream));
    const BlobRecord* R = ws->host_recs;
    // records by frame (counting sort of indices), then each frame ordered and grouped on its own: frames are
    // independent, so they are spread over host threads
    std::vector<unsigned>& first = ws->first; std::vector<unsigned>& order = ws->order;
    first.assign((size_t)n + 1, 0u); order.resize(nrec);
    for (unsigned i = 0; i < nrec; i++) first[R[i].frame + 1]++;
    for (int f = 0; f < n; f++) first[f + 1] += first[f];
    {
        std::vector<unsigned> at(first.begin(), first.end() - 1);
        for (unsigned i = 0; i < nrec; i++) order[at[R[i].frame]++] = i;
    }
    auto do_frames = [&](int f0, int f1)
    {
        for (int f = f0; f < f1; f++)
        {
            unsigned* o = order.data() + first[f];
            const int m = (int)(first[f + 1] - first[f]);
            if (!m) continue;
            std::sort(o, o + m, [&](unsigned a, unsigned b)
            {
                if (R[a].thr != R[b].thr) return R[a].thr < R[b].thr;
                return R[a].seq > R[b].seq;
            });
            counts_out[f] = group_frame(R, o, m, xy_out + (size_t)f * 2 * max_points, max_points);
        }
    };
    const int nthreads = std::max(1, std::min({ n / 4, 16, (int)std::thread::hardware_concurrency() }));
    if (nthreads <= 1) do_frames(0, n);
    else
    {
        std::vector<std::thread> pool;
        for (int t = 0; t < nthreads; t++)
            pool.emplace_back(do_frames, (int)((long long)n * t / nthreads), (int)((long long)n * (t + 1) / nthreads));
        for (auto& th : pool) th.join();
    }
    return 0;
}

// One chunk, synchronously. A single frame whose scratch overflows is run again with four times the space; a chunk
// of several frames that overflows is handed back to the caller (return 1), who runs its frames one by one (growing
// the scratch for a whole chunk could ask for tens of GB because of one busy frame).
int blob_find_frames(BlobWorkspace* ws, const FrameSet& fs, int32_t* xy_out, int32_t* counts_out, int max_points,
                     cudaStream_t stream, float* ms_out)
{
    if (ms_out) *ms_out = 0;
    for (int attempt = 0; ; attempt++)
    {
        if (blob_enqueue(ws, fs, stream)) return -1;
        const int rc = blob_finish(ws, xy_out, counts_out, max_points, ms_out);
        if (rc <= 0) return rc;
        if (fs.nframes > 1) return 1;
        if (attempt >= 6) { fprintf(stderr, "%s:%d in %s(): blob scratch still overflows after growing it 4096x. Sorry.\n", __FILE__, __LINE__, __func__); return -1; }
        ws->pts_per_frame = ws->pts_per_frame >= (1u << 29) ? ws->pts_per_frame : ws->pts_per_frame * 4; ws->rec_per_job *= 4;
        ws->queue_per_frame = ws->queue_per_frame >= (1u << 28) ? ws->queue_per_frame : ws->queue_per_frame * 4;
    }
}

}
