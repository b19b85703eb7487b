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
//   B2a blob_scan_kernel       segment starts by word operations: candidate first pixels of components / holes,
//                              and cut states on every 64th row / column (borders are followed in segments, so
//                              that a frame-sized outline does not keep one lane busy for milliseconds)
//   B2b blob_segment_kernel    one lane per segment start, lanes refilling from the queue: walks to the next
//                              segment start; length, area sum, earliest discovery position of its states
//   B2c blob_link_kernel, blob_chain_kernel   segments find their neighbours (one hash look-up each); one lane per
//                              candidate then follows the chain of segments in both directions
//                              until the two ends meet (the candidate is where OpenCV's scan discovers that
//                              border: a record with the border's length and exact area is emitted if the area
//                              passes the filter) or meets a segment the scan reaches earlier (dropped).
//                              No marks, no per-plane sequential pass.
//   B2d blob_points_kernel     one lane per kept border walks it once more: stores its points and sums the
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
    unsigned queue_cap;                         // segment starts of the whole chunk (queue entries, segments, candidate list)
    unsigned hash_mask;                         // slots of the state -> segment table, minus 1 (a power of two >= 2 * queue_cap)
    size_t plane_words;                         // storage words of one plane: (h + 2) rows of wpr words
    int origin;                                 // word offset of pixel (0,0) in a plane's storage
};
__device__ __forceinline__ const uint32_t* plane_ptr(const BlobGeom& g, const uint32_t* planes, int job)
{
    return planes + (size_t)job * g.plane_words + g.origin;
}

// device counters (uint32 each)
enum { kCntRecords = 0, kCntStatus = 1, kCntPoints = 2, kCntQueue = 3, kCntPointJob = 4, kCntQueueHead = 5, kCntSegs = 6, kCntCands = 7, kCntWords = 8 };

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
// B2a-c: segment starts (blob_scan_kernel), segments (blob_segment_kernel), chains (blob_chain_kernel); see blob_walk.cuh
// ------------------------------------------------------------------------------------------------
constexpr int kTileRows  = 16;                  // a tile = 32 words (1024 pixels) x 16 rows of one plane, one warp
constexpr unsigned kFull = 0xffffffffu;
// queue entry: job << 40 | kind << 32 | y << 16 | x   (kind 0: first pixel of a component, 1: first pixel of a hole,
// 2..5: cut state with background to the W / E / N / S -- the segment starts of blob_walk.cuh)

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
        uint32_t c[kTileRows + 2], el[kTileRows + 1], er[kTileRows + 1];
#pragma unroll
        for (int r = 0; r <= kTileRows + 1; r++) c[r] = (in && r <= rows + 1) ? B[(y0 + r - 1) * g.wpr + wd] : 0u;     // rows y0-1 .. y0+rows
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
        // per row: candidates (kinds 0, 1) and cut states (kinds 2..5: background to the W / E / N / S)
        uint32_t mk[kTileRows][6];
        unsigned cnt = 0;
#pragma unroll
        for (int r = 0; r < kTileRows; r++)
        {
            candidate_masks(c[r + 1], el[r + 1], c[r], el[r], er[r], &mk[r][0], &mk[r][1]);
            cut_masks(c[r + 1], el[r + 1], er[r + 1], c[r], c[r + 2], wd, y0 + r, &mk[r][2], &mk[r][3], &mk[r][4], &mk[r][5]);
            mk[r][2] &= ~mk[r][0];                              // (the same state: a component's first pixel on a cut row)
#pragma unroll
            for (int q = 0; q < 6; q++) { if (r >= rows) mk[r][q] = 0; cnt += __popc(mk[r][q]); }
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
                const unsigned long long hi = (unsigned long long)job << 40 | (unsigned long long)(y0 + r) << 16 | (unsigned)(wd * 32);
#pragma unroll
                for (int q = 0; q < 6; q++)
                {
                    uint32_t m = mk[r][q];
                    while (m) { const int bb = __ffs(m) - 1; m &= m - 1; if (at < g.queue_cap) queue[at] = hi | (unsigned long long)q << 32 | (unsigned)bb; at++; }
                }
            }
        }
    }
}

// One segment of a border: from a segment start to the next one
struct BlobSegment
{
    unsigned long long start, end;               // blobwalk::state_key of its first state / of the next segment's
    long long a00, a10, a01;                     // sums over its edges P->Q of d = Px*Qy - Qx*Py, d*(Px+Qx), d*(Py+Qy)
    int n;                                       // its states
    int pad;
};
// What the chains look at first lives in four plain arrays of queue_cap words each (one 4-byte load per hop for the many
// candidates that are dropped after a hop or two, instead of a 48-byte record): smallest discovery position of the
// segment's states (INT_MAX: none); the discovery position of its first state if that is a candidate start, else -1;
// the segment that starts where it ends; the segment that ends where it starts (blob_link_kernel).
struct SegMeta { int* min_disc; int* cand_pos; unsigned* next; unsigned* prev; };
__host__ __device__ inline SegMeta seg_meta(void* base, unsigned cap)
{
    SegMeta m; m.min_disc = (int*)base; m.cand_pos = m.min_disc + cap; m.next = (unsigned*)(m.cand_pos + cap); m.prev = m.next + cap;
    return m;
}
constexpr unsigned long long kNoKey = ~0ull;
__device__ __forceinline__ unsigned hash_of(unsigned long long key, unsigned mask)
{
    key ^= key >> 29; key *= 0x9E3779B97F4A7C15ull; key ^= key >> 32;
    return (unsigned)key & mask;
}

// B2b: every segment start is walked forward to the next one, by its own lane; lanes refill from the queue
__global__ void __launch_bounds__(128)
blob_segment_kernel(BlobGeom g, const uint32_t* __restrict__ planes, const unsigned long long* __restrict__ queue,
                    unsigned long long* __restrict__ hkeys, unsigned* __restrict__ hvals, BlobSegment* __restrict__ segs,
                    void* __restrict__ meta_base, unsigned* __restrict__ counters)
{
    const int lane = threadIdx.x & 31;
    const unsigned qn_all = counters[kCntQueue];
    if (qn_all > g.queue_cap) { if (threadIdx.x == 0 && blockIdx.x == 0) atomicExch(&counters[kCntStatus], 1u); return; }
    const unsigned qn = qn_all;
    const int max_steps = (int)min(4LL * g.w * g.h + 16, 0x7ffffff0LL);

    bool active = false, more = true;
    PlaneRef P; P.B = planes; P.w = g.w; P.h = g.h; P.wpr = g.wpr;
    BitWindow F;
    int x = 0, y = 0, k = 0, n = 0, min_disc = 0, cand_pos = -1;
    unsigned seg = 0, myslot = 0; unsigned long long skey = 0;
    bool fresh = false;
    long long a00 = 0, a10 = 0, a01 = 0;
    for (;;)
    {
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
                const int ex = (int)(e & 0xFFFFu), ey = (int)((e >> 16) & 0xFFFFu), kind = (int)((e >> 32) & 7u), job = (int)(e >> 40);
                P.B = plane_ptr(g, planes, job);
                F.init();
                bool ok = true;
                if (kind == 1) { x = ex - 1; y = ey; k = 1; }
                else { x = ex; y = ey; ok = state_after(P, F, ex, ey, kind == 3 ? 0 : kind == 4 ? 2 : kind == 5 ? 6 : 4, &k); }    // false: isolated pixel, area 0
                if (ok)
                {
                    // the state's slot in the table: whoever gets there first owns the segment (a state can be queued twice:
                    // a row cut that is also a column cut, ...)
                    skey = state_key(job, x, y, k);
                    unsigned slot = hash_of(skey, g.hash_mask);
                    for (;;)
                    {
                        const unsigned long long old = atomicCAS(&hkeys[slot], kNoKey, skey);
                        if (old == kNoKey) break;
                        if (old == skey) { ok = false; break; }
                        slot = (slot + 1) & g.hash_mask;
                    }
                    if (ok)
                    {
                        myslot = slot;
                        int disc; bool st;
                        classify_state(P, F, x, y, k, &disc, &st, &cand_pos);
                        min_disc = disc < 0 ? INT_MAX : disc;
                        n = 0; a00 = a10 = a01 = 0; active = true; fresh = true;
                    }
                }
            }
            // segment indices for the lanes that just started one: one atomic per warp (<= queue entries <= queue_cap)
            const unsigned starting = __ballot_sync(kFull, fresh);
            if (starting)
            {
                unsigned first = 0;
                if (lane == 0) first = atomicAdd(&counters[kCntSegs], (unsigned)__popc(starting));
                first = __shfl_sync(kFull, first, 0);
                if (fresh) { seg = first + __popc(starting & ((1u << lane) - 1u)); hvals[myslot] = seg; fresh = false; }
            }
        }
        if (!__any_sync(kFull, active)) { if (!more) break; continue; }
#pragma unroll 1
        for (int it = 0; it < 16; it++)
        {
            if (active)
            {
                const int px = x, py = y;
                int disc; bool st;
                step_fwd_ex(P, F, x, y, k, &disc, &st);
                { const long long d = (long long)(px * y - x * py); a00 += d; a10 += d * (px + x); a01 += d * (py + y); }
                n++;
                if (st || n > max_steps)
                {
                    if (!st) atomicExch(&counters[kCntStatus], 2u);         // cannot happen
                    BlobSegment s;
                    s.start = skey; s.end = state_key((int)(skey >> 33), x, y, k); s.a00 = a00; s.a10 = a10; s.a01 = a01; s.n = n; s.pad = 0;
                    segs[seg] = s;
                    const SegMeta M = seg_meta(meta_base, g.queue_cap);
                    M.min_disc[seg] = min_disc; M.cand_pos[seg] = cand_pos;
                    active = false;
                }
                else if (disc >= 0 && disc < min_disc) min_disc = disc;
            }
        }
    }
}

// B2c: segments find their neighbours: one hash look-up each, after which the chains are followed by index
__global__ void __launch_bounds__(256)
blob_link_kernel(BlobGeom g, const unsigned long long* __restrict__ hkeys, const unsigned* __restrict__ hvals,
                 const BlobSegment* __restrict__ segs, void* __restrict__ meta_base, unsigned* __restrict__ counters)
{
    if (counters[kCntStatus]) return;
    const unsigned nseg = counters[kCntSegs];
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < nseg; i += gridDim.x * blockDim.x)
    {
        const unsigned long long end = segs[i].end;
        unsigned slot = hash_of(end, g.hash_mask);
        bool found = true;
        while (hkeys[slot] != end)
        {
            if (hkeys[slot] == kNoKey) { found = false; break; }
            slot = (slot + 1) & g.hash_mask;
        }
        if (!found) { atomicExch(&counters[kCntStatus], 2u); continue; }                 // cannot happen: every segment ends at a segment start
        const unsigned j = hvals[slot];
        const SegMeta M = seg_meta(meta_base, g.queue_cap);
        M.next[i] = j;
        M.prev[j] = i;
    }
}

// B2d: every candidate follows the chain of segments from its own, in both directions at once (as the walkers of
// single states did before: a candidate that is not its border's start meets a segment holding a state the raster
// scan reaches earlier within a few hops one way or the other, and is dropped). If the two ends meet instead, the
// candidate is where the scan discovers the border, whose length and area are then known -- a record is emitted if
// the area passes the filter.
__global__ void __launch_bounds__(128)
blob_chain_kernel(BlobGeom g, const uint32_t* __restrict__ planes, const BlobSegment* __restrict__ segs, void* __restrict__ meta_base,
                  BlobRecord* __restrict__ recs, unsigned* __restrict__ counters)
{
    if (counters[kCntStatus]) return;
    const unsigned nseg = counters[kCntSegs];
    const SegMeta M = seg_meta(meta_base, g.queue_cap);
    for (unsigned c0 = blockIdx.x * blockDim.x + threadIdx.x; c0 < nseg; c0 += gridDim.x * blockDim.x)
    {
        const int pos = M.cand_pos[c0];
        if (pos < 0) continue;                   // (a cut state: not where any border can be discovered)
        if (M.min_disc[c0] < pos) continue;
        const BlobSegment c = segs[c0];
        long long a00 = c.a00, a10 = c.a10, a01 = c.a01, n = c.n;
        unsigned f = c0, b = c0;                 // covered so far: the segments from b forward to f (through c0)
        bool keep = true;
        for (unsigned hops = 0; ; hops++)
        {
            if (hops > g.queue_cap) { atomicExch(&counters[kCntStatus], 2u); keep = false; break; }    // cannot happen
            const unsigned fn = M.next[f];
            if (fn == b) break;
            if (M.min_disc[fn] < pos) { keep = false; break; }
            { const BlobSegment& s = segs[fn]; a00 += s.a00; a10 += s.a10; a01 += s.a01; n += s.n; }
            f = fn;
            const unsigned bp = M.prev[b];
            if (bp == f) break;
            if (M.min_disc[bp] < pos) { keep = false; break; }
            { const BlobSegment& s = segs[bp]; a00 += s.a00; a10 += s.a10; a01 += s.a01; n += s.n; }
            b = bp;
        }
        if (!keep) continue;
        // filterByArea: m00 = |a00| / 2 in [20, 80000) -- exact in integers. Everything else is dropped here.
        const long long aa = a00 < 0 ? -a00 : a00;
        if (aa < 40 || aa >= 160000) continue;
        const int job = (int)(c.start >> 33);
        {
            // filterByColor already here (the centre needs only the first-order sums, which the segments carry): a border
            // whose centre pixel is foreground is rejected whatever else is true of it -- on board frames that is half
            // of the borders the area filter keeps (the white regions) -- so its points are never stored
            const double sgn = a00 > 0 ? 1.0 : -1.0;
            const double m00 = __dmul_rn((double)a00, sgn * 0.5);
            const double m10 = __dmul_rn((double)a10, sgn * 0.16666666666666666666666666666667);
            const double m01 = __dmul_rn((double)a01, sgn * 0.16666666666666666666666666666667);
            const int rx = __double2int_rn(__ddiv_rn(m10, m00)), ry = __double2int_rn(__ddiv_rn(m01, m00));
            if (rx < 0 || rx >= g.w || ry < 0 || ry >= g.h) continue;
            const uint32_t* B = plane_ptr(g, planes, job);
            if ((B[ry * g.wpr + (rx >> 5)] >> (rx & 31)) & 1u) continue;
        }
        const unsigned idx = atomicAdd(&counters[kCntRecords], 1u);
        const unsigned off = atomicAdd(&counters[kCntPoints], (unsigned)n);
        if (idx >= g.rec_cap || off > g.pts_cap || (unsigned long long)n > g.pts_cap - off) { atomicExch(&counters[kCntStatus], 1u); continue; }
        BlobRecord r;
        r.frame = job / kNThr; r.thr = job % kNThr; r.seq = pos; r.n = (int)n; r.pts_off = off;
        r.sx = (int)((c.start >> 3) & 0x7FFFu); r.sy = (int)((c.start >> 18) & 0x7FFFu); r.sk = (int)(c.start & 7u);
        r.xmin = r.xmax = r.ymin = r.ymax = 0;
        r.a00 = a00; r.a10 = r.a01 = r.a20 = r.a11 = r.a02 = 0;
        r.hull2 = 0; r.cx = r.cy = r.radius = 0; r.colour_ok = 0; r.big = 0;
        recs[idx] = r;
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
        __syncwarp();                            // (the previous border's readers of the warp's slice are done)
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
    void* hkeys = nullptr; void* hvals = nullptr; void* segs = nullptr; void* segmeta = nullptr;
    size_t planes_b = 0, pts_b = 0, recs_b = 0, scratch_b = 0, queue_b = 0, hkeys_b = 0, hvals_b = 0, segs_b = 0, segmeta_b = 0;
    unsigned pts_per_frame = 1u << 21, rec_per_job = 512, queue_per_frame = 1u << 17;
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
    cudaFree(ws->hkeys); cudaFree(ws->hvals); cudaFree(ws->segs); cudaFree(ws->segmeta);
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

void blob_workspace_reset_capacity(BlobWorkspace* ws) { ws->pts_per_frame = 1u << 21; ws->rec_per_job = 512; ws->queue_per_frame = 1u << 17; }

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
    unsigned slots = 1024; while (slots < 2ull * g.queue_cap && slots < 0x80000000u) slots <<= 1;
    g.hash_mask = slots - 1;
    BLOB_TRY(grow(&ws->hkeys, &ws->hkeys_b, (size_t)slots * 8));
    BLOB_TRY(grow(&ws->hvals, &ws->hvals_b, (size_t)slots * 4));
    BLOB_TRY(grow(&ws->segs, &ws->segs_b, (size_t)g.queue_cap * sizeof(BlobSegment)));
    BLOB_TRY(grow(&ws->segmeta, &ws->segmeta_b, (size_t)g.queue_cap * 16));
    BLOB_TRY(grow(&ws->pts, &ws->pts_b, (size_t)g.pts_cap * 4));
    BLOB_TRY(grow(&ws->recs, &ws->recs_b, (size_t)g.rec_cap * sizeof(BlobRecord)));
    unsigned* counters = (unsigned*)ws->counters;
    BLOB_TRY(cudaMemsetAsync(ws->counters, 0, kCntWords * 4, stream));
    BLOB_TRY(cudaEventRecord(ws->e0, stream));
    BLOB_TRY(cudaMemsetAsync(ws->hkeys, 0xFF, ((size_t)g.hash_mask + 1) * 8, stream));
    blob_binarize_kernel<<<dim3((g.wpr + 127) / 128, plane_rows(g.h), n), 128, 0, stream>>>(fs, g, (uint32_t*)ws->planes, aligned);
    blob_scan_kernel<<<148 * 8, 256, 0, stream>>>(g, (const uint32_t*)ws->planes, (unsigned long long*)ws->queue, counters,
                                                  (unsigned)ntiles64, nstrips, ncb);
    blob_segment_kernel<<<148 * 8, 128, 0, stream>>>(g, (const uint32_t*)ws->planes, (const unsigned long long*)ws->queue,
                                                     (unsigned long long*)ws->hkeys, (unsigned*)ws->hvals, (BlobSegment*)ws->segs, ws->segmeta, counters);
    blob_link_kernel<<<148 * 8, 256, 0, stream>>>(g, (const unsigned long long*)ws->hkeys, (const unsigned*)ws->hvals, (const BlobSegment*)ws->segs, ws->segmeta, counters);
    blob_chain_kernel<<<148 * 8, 128, 0, stream>>>(g, (const uint32_t*)ws->planes, (const BlobSegment*)ws->segs, ws->segmeta, (BlobRecord*)ws->recs, counters);
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
