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
// GPU decomposition (all 17 thresholds of all frames of a chunk in flight at once):
//   B1 blob_binarize_kernel  gray -> 17 bit planes per frame (one pass over the frame, 1 B/px read)
//   B2 blob_trace_kernel     one thread per (frame, threshold) replays OpenCV's raster scan and
//                            border following on the bit plane. The scan visits only horizontal 0/1
//                            transitions (32 pixels per word op); the pixel states of the original
//                            (1 = untouched, 2 = visited, -126 = visited + "east neighbour examined
//                            and zero") live in two more bit planes. While following a border it
//                            accumulates the polygon moments as exact 64-bit integers (Green's
//                            theorem terms are integers; order does not matter) and stores the
//                            points; borders whose area is outside [20, 80000) are dropped at once.
//   B3 blob_contour_kernel   one CTA per surviving border: convex-hull area from per-column
//                            extremes (exact integers), colour test at the rounded centre, median
//                            point distance by radix selection on the IEEE bit patterns.
//   host                     the remaining double arithmetic (inertia, convexity ratio) and the
//                            grouping across thresholds, a few hundred centres per frame, in the
//                            reference's operation order (no FMA contraction: see build.py).
#include <cuda_runtime.h>
#include <float.h>
#include <math.h>
#include <stdio.h>
#include <string.h>
#include <algorithm>
#include <vector>

#include "kernels.cuh"

namespace mrgb200
{
namespace
{
constexpr int kNThr = 17;                       // thresholds 50, 60, ..., 210 (min 50, max 220, step 10)
__host__ __device__ inline int thr_value(int k) { return 50 + 10 * k; }

struct BlobRecord
{
    int frame, thr, seq, n;                     // seq = discovery order among the survivors of (frame, thr)
    unsigned pts_off;                           // first point, within the (frame, thr) point region
    int xmin, xmax, ymin, ymax;
    long long a00, a10, a01, a20, a11, a02;     // Green's-theorem sums over the directed border edges
    long long hull2;                            // twice the convex hull's area
    double cx, cy, radius;
    int colour_ok, pad;
};

struct BlobGeom
{
    int w, h, wpr;                              // pixels, 32-bit words per bit-plane row
    int nframes;
    unsigned pts_cap;                           // points per (frame, thr) region
    unsigned rec_cap;
};

// ------------------------------------------------------------------------------------------------
// B1: bit planes. plane(f,k)[y][wd] bit b = gray(f, y, 32*wd + b) > thr_k
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
blob_binarize_kernel(FrameSet fs, BlobGeom g, uint32_t* __restrict__ planes)
{
    const int wd = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, f = blockIdx.z;
    if (wd >= g.wpr) return;
    const uint8_t* row = fs.base + (size_t)f * fs.frame_stride + (size_t)y * fs.pitch;
    uint8_t v[32];
    const int x0 = wd * 32;
#pragma unroll
    for (int i = 0; i < 32; i++) v[i] = x0 + i < g.w ? row[x0 + i] : 0;
    for (int k = 0; k < kNThr; k++)
    {
        const int t = thr_value(k);
        uint32_t bits = 0;
#pragma unroll
        for (int i = 0; i < 32; i++) bits |= (uint32_t)(v[i] > t) << i;
        planes[(((size_t)f * kNThr + k) * g.h + y) * g.wpr + wd] = bits;
    }
}

// ------------------------------------------------------------------------------------------------
// B2: raster scan + border following, one thread per (frame, threshold)
// ------------------------------------------------------------------------------------------------
__constant__ int kDx[8] = { 1, 1, 0, -1, -1, -1, 0, 1 };     // 0 = east, then counter-clockwise on the screen
__constant__ int kDy[8] = { 0, -1, -1, -1, 0, 1, 1, 1 };

// One thread's view of its bit planes. The thread keeps, in registers, the 3 x 3 words of B around
// its position (rows cy-1..cy+1, words cwd-1..cwd+1) and the V/R words it is marking (written back
// when it moves to another word): while a border is followed most steps need no load at all, a
// vertical step needs one round of three independent loads.
struct Plane
{
    const uint32_t* B; uint32_t* V; uint32_t* R;
    int w, h, wpr;
    int cy, cwd; uint32_t b[3][3];
    int my, mwd; uint32_t vword, rword; bool dirty;

    __device__ __forceinline__ void init() { cy = -0x40000000; cwd = -2; my = -1; mwd = -1; vword = rword = 0; dirty = false; }
    __device__ __forceinline__ void load_row(int slot, int y)
    {
#pragma unroll
        for (int c = 0; c < 3; c++)
        {
            const int wc = cwd - 1 + c;
            b[slot][c] = ((unsigned)y < (unsigned)h && (unsigned)wc < (unsigned)wpr) ? B[(size_t)y * wpr + wc] : 0u;
        }
    }
    __device__ __forceinline__ void seek(int x, int y)
    {
        const int wd = x >> 5;
        if (wd == cwd && y == cy) return;
        if (wd == cwd && y == cy + 1)
        {
#pragma unroll
            for (int c = 0; c < 3; c++) { b[0][c] = b[1][c]; b[1][c] = b[2][c]; }
            cy = y; load_row(2, y + 1);
        }
        else if (wd == cwd && y == cy - 1)
        {
#pragma unroll
            for (int c = 0; c < 3; c++) { b[2][c] = b[1][c]; b[1][c] = b[0][c]; }
            cy = y; load_row(0, y - 1);
        }
        else
        {
            cwd = wd; cy = y;
            load_row(0, y - 1); load_row(1, y); load_row(2, y + 1);
        }
    }
    // bits (x-1, x, x+1) of a cached row as bits 0..2
    __device__ __forceinline__ uint32_t row3(int slot, int bpos) const
    {
        const unsigned long long w64 = ((unsigned long long)b[slot][1] << 32) | b[slot][0];
        uint32_t v = (uint32_t)(w64 >> (31 + bpos)) & 7u;
        if (bpos == 31) v |= (b[slot][2] & 1u) << 2;
        return v;
    }
    // bit d = the neighbour of (x, y) in direction d (0 = E, 1 = NE, 2 = N, 3 = NW, 4 = W, 5 = SW, 6 = S, 7 = SE) is foreground
    __device__ __forceinline__ uint32_t nbr8(int x, int y)
    {
        seek(x, y);
        const int bpos = x & 31;
        const uint32_t up = row3(0, bpos), mid = row3(1, bpos), dn = row3(2, bpos);
        return ((mid >> 2) & 1u) | (((up >> 2) & 1u) << 1) | (((up >> 1) & 1u) << 2) | ((up & 1u) << 3) |
               ((mid & 1u) << 4) | ((dn & 1u) << 5) | (((dn >> 1) & 1u) << 6) | (((dn >> 2) & 1u) << 7);
    }
    // mark words: write-back cache of one V word and one R word
    __device__ __forceinline__ void flush_marks()
    {
        if (dirty) { V[(size_t)my * wpr + mwd] = vword; R[(size_t)my * wpr + mwd] = rword; dirty = false; }
    }
    __device__ __forceinline__ void seek_marks(int x, int y)
    {
        const int wd = x >> 5;
        if (wd == mwd && y == my) return;
        flush_marks();
        my = y; mwd = wd;
        vword = V[(size_t)y * wpr + wd]; rword = R[(size_t)y * wpr + wd];
    }
    __device__ __forceinline__ bool visited(int x, int y) { seek_marks(x, y); return (vword >> (x & 31)) & 1u; }
    __device__ __forceinline__ bool rflag(int x, int y)   { seek_marks(x, y); return (rword >> (x & 31)) & 1u; }
    __device__ __forceinline__ void set_visited(int x, int y) { seek_marks(x, y); vword |= 1u << (x & 31); dirty = true; }
    __device__ __forceinline__ void set_rflag(int x, int y)   { seek_marks(x, y); rword |= 1u << (x & 31); dirty = true; }
};

struct Tracer
{
    Plane P;
    uint32_t* pts; unsigned npts, pts_cap;
    BlobRecord* recs; unsigned* rec_count; unsigned rec_cap;
    int* status;
    int frame, thr, seq;

    // follow the border that starts at (x0, y0); false = out of space
    __device__ __forceinline__ bool trace(int x0, int y0, bool is_hole)
    {
        const unsigned start = npts;
        // Only the area term is accumulated here (it decides at once whether the border is kept); the
        // other moments and the bounding box are summed over the stored points, in parallel, by B3.
        long long a00 = 0;
        int fx = x0, fy = y0, px = x0, py = y0;          // first / previous emitted point
        int n = 0;
        auto emit = [&](int x, int y) -> bool
        {
            if (npts >= pts_cap) return false;
            pts[npts++] = (uint32_t)x | ((uint32_t)y << 16);
            a00 += px * y - x * py;                      // coordinates < 2^15: the products fit 32 bits
            px = x; py = y; n++;
            return true;
        };

        // neighbour search on an 8-bit mask of the 3x3 neighbourhood (bit d = direction d is
        // foreground): the three rows are fetched with independent loads, the rotation is bit math
        int s_end = is_hole ? 0 : 4, s, x1, y1;
        {
            // first neighbour clockwise from s_end: directions s_end-1, s_end-2, ..., s_end (mod 8)
            const uint32_t m = P.nbr8(x0, y0);
            const uint32_t rot = ((m | (m << 8)) >> s_end) & 0xFFu;          // bit j = direction s_end + j
            // clockwise order = j = 7, 6, ..., 1: the highest set bit (direction s_end itself, j = 0, is
            // background at every start the scan can produce); none = isolated pixel
            const int hb = rot ? 31 - __clz(rot) : 0;
            s = (s_end + hb) & 7;
            x1 = x0 + kDx[s]; y1 = y0 + kDy[s];
        }
        if (s == s_end)
        {
            P.set_visited(x0, y0); P.set_rflag(x0, y0);          // isolated pixel
            if (!emit(x0, y0)) return false;
        }
        else
        {
            int x3 = x0, y3 = y0;
            for (;;)
            {
                s_end = s;
                // first neighbour counter-clockwise from s+1: directions s+1, s+2, ... (mod 8)
                const uint32_t m = P.nbr8(x3, y3);
                const uint32_t rot = ((m | (m << 8)) >> ((s + 1) & 7)) & 0xFFu;   // bit j = direction s + 1 + j
                const int j = __ffs(rot) - 1;                                    // a neighbour always exists: we came from one
                const int s_raw = s + 1 + j;                                     // what the original's ++s loop ends on (<= 15)
                s = s_raw & 7;
                const int x4 = x3 + kDx[s], y4 = y3 + kDy[s];
                // marks of (x3, y3): one look-up of the cached mark words per step
                {
                    P.seek_marks(x3, y3);
                    const uint32_t bit = 1u << (x3 & 31);
                    if ((unsigned)(s - 1) < (unsigned)s_end) { P.vword |= bit; P.rword |= bit; P.dirty = true; }
                    else if (!(P.vword & bit)) { P.vword |= bit; P.dirty = true; }
                }
                if (!emit(x3, y3)) return false;
                if (x4 == x0 && y4 == y0 && x3 == x1 && y3 == y1) break;
                x3 = x4; y3 = y4;
                s = (s + 4) & 7;
            }
        }
        a00 += px * fy - fx * py;                        // closing edge: last point -> first point
        // filterByArea: m00 = |a00| / 2 in [20, 80000) -- exact in integers. Everything else is dropped here.
        const long long aa = a00 < 0 ? -a00 : a00;
        if (aa < 40 || aa >= 160000) { npts = start; return true; }
        const unsigned idx = atomicAdd(rec_count, 1u);
        if (idx >= rec_cap) return false;
        BlobRecord r;
        r.frame = frame; r.thr = thr; r.seq = seq++; r.n = n; r.pts_off = start;
        r.xmin = r.xmax = r.ymin = r.ymax = 0;
        r.a00 = a00; r.a10 = r.a01 = r.a20 = r.a11 = r.a02 = 0;
        r.hull2 = 0; r.cx = r.cy = r.radius = 0; r.colour_ok = 0; r.pad = 0;
        recs[idx] = r;
        return true;
    }
};

__global__ void __launch_bounds__(32)
blob_trace_kernel(BlobGeom g, const uint32_t* __restrict__ planes, uint32_t* marksV, uint32_t* marksR,
                  uint32_t* pts, BlobRecord* recs, unsigned* rec_count, int* status)
{
    // One warp per (frame, threshold). The raster scan for 0/1 transitions is done by all 32 lanes
    // (a lane looks at four words = 128 pixels; a 4K row is one coalesced 512-byte load); the groups that
    // hold transitions are then handed, in raster order, to lane 0, which owns the mark planes and
    // follows the borders exactly as the sequential original does.
    const int lane = threadIdx.x;
    const int job = blockIdx.x, f = job / kNThr, k = job % kNThr;
    const size_t poff = ((size_t)f * kNThr + k) * g.h * g.wpr;
    Tracer T;
    T.P.B = planes + poff; T.P.V = marksV + poff; T.P.R = marksR + poff;
    T.P.w = g.w; T.P.h = g.h; T.P.wpr = g.wpr; T.P.init();
    T.pts = pts + (size_t)job * g.pts_cap; T.npts = 0; T.pts_cap = g.pts_cap;
    T.recs = recs; T.rec_count = rec_count; T.rec_cap = g.rec_cap; T.status = status;
    T.frame = f; T.thr = k; T.seq = 0;

    const int groups = g.wpr / 4;                         // wpr is a multiple of 4, the planes are 16-byte aligned
    for (int y = 0; y < g.h; y++)
    {
        const uint4* brow = reinterpret_cast<const uint4*>(T.P.B + (size_t)y * g.wpr);
        uint32_t carry = 0;                               // last pixel of the previous 32 groups
        for (int g0 = 0; g0 < groups; g0 += 32)
        {
            const int gi = g0 + lane;
            const uint4 q = gi < groups ? brow[gi] : make_uint4(0, 0, 0, 0);
            const uint32_t pw = __shfl_up_sync(0xffffffffu, q.w, 1);
            const uint32_t prevbit = lane == 0 ? carry : pw >> 31;
            carry = __shfl_sync(0xffffffffu, q.w, 31) >> 31;
            // pixels that differ from their left neighbour
            uint32_t e0 = q.x ^ ((q.x << 1) | prevbit), e1 = q.y ^ ((q.y << 1) | (q.x >> 31));
            uint32_t e2 = q.z ^ ((q.z << 1) | (q.y >> 31)), e3 = q.w ^ ((q.w << 1) | (q.z >> 31));
            const int left = g.w - gi * 128;              // only x < w is examined (the scan stops before the frame)
            if (left < 128)
            {
                e0 &= left >= 32 ? ~0u : (left > 0 ? (1u << left) - 1 : 0u);
                e1 &= left >= 64 ? ~0u : (left > 32 ? (1u << (left - 32)) - 1 : 0u);
                e2 &= left >= 96 ? ~0u : (left > 64 ? (1u << (left - 64)) - 1 : 0u);
                e3 &= left > 96 ? (1u << (left - 96)) - 1 : 0u;
            }
            uint32_t m = __ballot_sync(0xffffffffu, (e0 | e1 | e2 | e3) != 0);
            while (m)
            {
                const int src = __ffs(m) - 1;
                m &= m - 1;
                const uint32_t cx = __shfl_sync(0xffffffffu, q.x, src), cy = __shfl_sync(0xffffffffu, q.y, src);
                const uint32_t cz = __shfl_sync(0xffffffffu, q.z, src), cw = __shfl_sync(0xffffffffu, q.w, src);
                const uint32_t f0 = __shfl_sync(0xffffffffu, e0, src), f1 = __shfl_sync(0xffffffffu, e1, src);
                const uint32_t f2 = __shfl_sync(0xffffffffu, e2, src), f3 = __shfl_sync(0xffffffffu, e3, src);
                int failed = 0;
                if (lane == 0)
                {
#pragma unroll 1
                    for (int kk = 0; kk < 4 && !failed; kk++)
                    {
                        uint32_t ev = kk == 0 ? f0 : kk == 1 ? f1 : kk == 2 ? f2 : f3;
                        const uint32_t cur = kk == 0 ? cx : kk == 1 ? cy : kk == 2 ? cz : cw;
                        while (ev)
                        {
                            const int bb = __ffs(ev) - 1;
                            ev &= ev - 1;
                            const int x = ((g0 + src) * 4 + kk) * 32 + bb;
                            // 0 -> 1: an outer border starts here unless the pixel was already visited;
                            // 1 -> 0: a hole border starts at x-1 unless that pixel carries the east flag
                            const bool hole = !((cur >> bb) & 1u);
                            const int sx = x - (hole ? 1 : 0);
                            const bool start = hole ? !T.P.rflag(sx, y) : !T.P.visited(sx, y);
                            if (start && !T.trace(sx, y, hole)) { failed = 1; break; }
                        }
                    }
                }
                failed = __shfl_sync(0xffffffffu, failed, 0);
                if (failed) { if (lane == 0) atomicExch(status, 1); return; }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// B3: hull area, colour test, median radius of one surviving border per CTA
// ------------------------------------------------------------------------------------------------
constexpr int kB3Threads = 128;

__global__ void __launch_bounds__(kB3Threads)
blob_contour_kernel(BlobGeom g, const uint32_t* __restrict__ planes, const uint32_t* __restrict__ pts,
                    BlobRecord* recs, const unsigned* __restrict__ rec_count, int* scratch, int scratch_stride)
{
    __shared__ unsigned hist[256];
    __shared__ unsigned long long sel_prefix;
    __shared__ unsigned sel_rank;
    __shared__ double s_cx, s_cy;
    __shared__ long long red_ll[kB3Threads / 32][5];
    __shared__ int red_i[kB3Threads / 32][4];
    __shared__ int s_need_radius;
    const unsigned nrec = min(*rec_count, g.rec_cap);
    int* ylo = scratch + (size_t)blockIdx.x * scratch_stride;       // per column of the bounding box
    int* yhi = ylo + g.w;
    int* stk = yhi + g.w;                                           // hull chain: (x, y) pairs
    for (unsigned ri = blockIdx.x; ri < nrec; ri += gridDim.x)
    {
        BlobRecord& r = recs[ri];
        const uint32_t* p = pts + ((size_t)r.frame * kNThr + r.thr) * g.pts_cap + r.pts_off;
        const int n = r.n;
        // Green's-theorem sums over the directed edges (point i-1 -> point i, cyclically) and the
        // bounding box: exact integers, any order
        {
            long long t10 = 0, t01 = 0, t20 = 0, t11 = 0, t02 = 0;
            int bx0 = INT_MAX, bx1 = INT_MIN, by0 = INT_MAX, by1 = INT_MIN;
            for (int i = threadIdx.x; i < n; i += kB3Threads)
            {
                const uint32_t q = p[i], qp = p[i == 0 ? n - 1 : i - 1];
                const long long x = q & 0xFFFF, y = q >> 16, xp = qp & 0xFFFF, yp = qp >> 16;
                const long long dxy = xp * y - x * yp, xs = xp + x, ys2 = yp + y;
                t10 += dxy * xs; t01 += dxy * ys2;
                t20 += dxy * (xp * xs + x * x);
                t11 += dxy * (xp * (ys2 + yp) + x * (ys2 + y));
                t02 += dxy * (yp * ys2 + y * y);
                bx0 = min(bx0, (int)x); bx1 = max(bx1, (int)x); by0 = min(by0, (int)y); by1 = max(by1, (int)y);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1)
            {
                t10 += __shfl_down_sync(0xffffffffu, t10, o); t01 += __shfl_down_sync(0xffffffffu, t01, o);
                t20 += __shfl_down_sync(0xffffffffu, t20, o); t11 += __shfl_down_sync(0xffffffffu, t11, o);
                t02 += __shfl_down_sync(0xffffffffu, t02, o);
                bx0 = min(bx0, __shfl_down_sync(0xffffffffu, bx0, o)); bx1 = max(bx1, __shfl_down_sync(0xffffffffu, bx1, o));
                by0 = min(by0, __shfl_down_sync(0xffffffffu, by0, o)); by1 = max(by1, __shfl_down_sync(0xffffffffu, by1, o));
            }
            if ((threadIdx.x & 31) == 0)
            {
                const int wi = threadIdx.x >> 5;
                red_ll[wi][0] = t10; red_ll[wi][1] = t01; red_ll[wi][2] = t20; red_ll[wi][3] = t11; red_ll[wi][4] = t02;
                red_i[wi][0] = bx0; red_i[wi][1] = bx1; red_i[wi][2] = by0; red_i[wi][3] = by1;
            }
            __syncthreads();
            if (threadIdx.x == 0)
            {
                long long u[5] = { 0, 0, 0, 0, 0 };
                int c0 = INT_MAX, c1 = INT_MIN, c2 = INT_MAX, c3 = INT_MIN;
                for (int wq = 0; wq < kB3Threads / 32; wq++)
                {
                    for (int k = 0; k < 5; k++) u[k] += red_ll[wq][k];
                    c0 = min(c0, red_i[wq][0]); c1 = max(c1, red_i[wq][1]); c2 = min(c2, red_i[wq][2]); c3 = max(c3, red_i[wq][3]);
                }
                r.a10 = u[0]; r.a01 = u[1]; r.a20 = u[2]; r.a11 = u[3]; r.a02 = u[4];
                r.xmin = c0; r.xmax = c1; r.ymin = c2; r.ymax = c3;
            }
            __syncthreads();
        }
        const int bw = r.xmax - r.xmin + 1;
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
            // convex hull of the border = hull of the per-column extremes (every column of the box holds
            // a border point). Monotone chains over the columns; the area comes out as an exact integer.
            auto chain_sum = [&](const int* ys, bool lower) -> long long
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
            };
            // polygon: lower chain left -> right, up the last column, upper chain right -> left, down the first
            const long long sl = chain_sum(ylo, true), su = chain_sum(yhi, false);
            const long long xe = bw - 1;
            long long a2 = sl + xe * ((long long)yhi[bw-1] - ylo[bw-1]) - su;      // first column: x = 0 contributes nothing
            r.hull2 = a2 < 0 ? -a2 : a2;

            // centre, with the rounding of cv::moments / SimpleBlobDetector: m = a * (+-1/2, +-1/6), c = m10 / m00
            const double sgn = r.a00 > 0 ? 1.0 : -1.0;
            const double m00 = __dmul_rn((double)r.a00, sgn * 0.5);
            const double m10 = __dmul_rn((double)r.a10, sgn * 0.16666666666666666666666666666667);
            const double m01 = __dmul_rn((double)r.a01, sgn * 0.16666666666666666666666666666667);
            const double cx = __ddiv_rn(m10, m00), cy = __ddiv_rn(m01, m00);
            r.cx = cx; r.cy = cy; s_cx = cx; s_cy = cy;
            // filterByColor: the binary image must be 0 at (cvRound(cy), cvRound(cx))
            const int rx = __double2int_rn(cx), ry = __double2int_rn(cy);
            int ok = 0;
            if (rx >= 0 && rx < g.w && ry >= 0 && ry < g.h)
            {
                const uint32_t* B = planes + ((size_t)r.frame * kNThr + r.thr) * g.h * g.wpr;
                ok = !((B[(size_t)ry * g.wpr + (rx >> 5)] >> (rx & 31)) & 1u);
            }
            r.colour_ok = ok;
            // The host rejects the record if the colour test fails or area / hullArea < 0.95f. Where that is
            // certain (colour, or a ratio below 0.94: far from any rounding question) the median is not needed.
            const long long aa = r.a00 < 0 ? -r.a00 : r.a00;
            s_need_radius = ok && aa * 100 >= r.hull2 * 94;
        }
        __syncthreads();
        if (!s_need_radius) { __syncthreads(); continue; }
        // median of the point distances to the centre: order statistics (n-1)/2 and n/2 of
        // d2 = dx*dx + dy*dy, selected on the bit patterns (non-negative doubles order like integers)
        const double cx = s_cx, cy = s_cy;
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
                    const uint32_t q = p[i];
                    const double dx = __dsub_rn(cx, (double)(int)(q & 0xFFFF)), dy = __dsub_rn(cy, (double)(int)(q >> 16));
                    const unsigned long long key = (unsigned long long)__double_as_longlong(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
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
// recs: this frame's records sorted by (thr ascending, seq DESCENDING): OpenCV hands contours over
// in reverse discovery order.
int group_frame(const BlobRecord* recs, int nrec, int32_t* xy_out, int max_points)
{
    std::vector<std::vector<Center>> centers;
    int i = 0;
    while (i < nrec)
    {
        const int thr = recs[i].thr;
        std::vector<std::vector<Center>> fresh;
        for (; i < nrec && recs[i].thr == thr; i++)
        {
            Center c;
            if (!center_from_record(recs[i], &c)) continue;
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
    void* planes = nullptr; void* marksV = nullptr; void* marksR = nullptr; void* pts = nullptr; void* recs = nullptr;
    void* scratch = nullptr; void* counters = nullptr;
    size_t planes_b = 0, marks_b = 0, pts_b = 0, recs_b = 0, scratch_b = 0;
    unsigned pts_cap = 1u << 18, rec_per_job = 512;
    std::vector<BlobRecord> host_recs;
};

BlobWorkspace* blob_workspace_create() { return new BlobWorkspace(); }
void blob_workspace_destroy(BlobWorkspace* ws)
{
    if (!ws) return;
    cudaFree(ws->planes); cudaFree(ws->marksV); cudaFree(ws->marksR); cudaFree(ws->pts); cudaFree(ws->recs);
    cudaFree(ws->scratch); cudaFree(ws->counters);
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

// Blob detection over device-resident frames. xy_out: HOST [nframes][max_points][2] int32 (scaled
// by 1000), counts_out: HOST [nframes]. Synchronous. Returns 0, -1 on a CUDA failure, or 1 when the
// scratch of a multi-frame chunk overflowed (nothing was produced: run the frames one at a time).
void blob_workspace_reset_capacity(BlobWorkspace* ws) { ws->pts_cap = 1u << 18; ws->rec_per_job = 512; }

int blob_find_frames(BlobWorkspace* ws, const FrameSet& fs, int32_t* xy_out, int32_t* counts_out, int max_points,
                     cudaStream_t stream, float* ms_out)
{
    const int n = fs.nframes;
    for (int i = 0; i < n; i++) counts_out[i] = 0;
    if (n <= 0 || fs.w <= 0 || fs.h <= 0) return 0;
    BlobGeom g;
    g.w = fs.w; g.h = fs.h; g.wpr = ((fs.w + 31) / 32 + 3) & ~3; g.nframes = n;      // rows of the planes: multiples of 16 bytes
    const size_t plane_words = (size_t)n * kNThr * g.h * g.wpr;
    BLOB_TRY(grow(&ws->planes, &ws->planes_b, plane_words * 4));
    {
        size_t have = ws->marks_b;
        BLOB_TRY(grow(&ws->marksV, &have, plane_words * 4));
        BLOB_TRY(grow(&ws->marksR, &ws->marks_b, plane_words * 4));
    }
    if (!ws->counters) BLOB_TRY(cudaMalloc(&ws->counters, 16));
    const int b3_blocks = 148 * 4;
    const int scratch_stride = 4 * fs.w + 8;
    BLOB_TRY(grow(&ws->scratch, &ws->scratch_b, (size_t)b3_blocks * scratch_stride * sizeof(int)));

    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (ms_out) { cudaEventCreate(&e0); cudaEventCreate(&e1); *ms_out = 0; }
    for (int attempt = 0; ; attempt++)
    {
        g.pts_cap = ws->pts_cap; g.rec_cap = ws->rec_per_job * (unsigned)n * kNThr;
        BLOB_TRY(grow(&ws->pts, &ws->pts_b, (size_t)n * kNThr * g.pts_cap * 4));
        BLOB_TRY(grow(&ws->recs, &ws->recs_b, (size_t)g.rec_cap * sizeof(BlobRecord)));
        unsigned* rec_count = (unsigned*)ws->counters; int* status = (int*)ws->counters + 1;
        BLOB_TRY(cudaMemsetAsync(ws->counters, 0, 16, stream));
        BLOB_TRY(cudaMemsetAsync(ws->marksV, 0, plane_words * 4, stream));
        BLOB_TRY(cudaMemsetAsync(ws->marksR, 0, plane_words * 4, stream));
        if (e0) cudaEventRecord(e0, stream);
        blob_binarize_kernel<<<dim3((g.wpr + 127) / 128, g.h, n), 128, 0, stream>>>(fs, g, (uint32_t*)ws->planes);
        blob_trace_kernel<<<n * kNThr, 32, 0, stream>>>(g, (const uint32_t*)ws->planes, (uint32_t*)ws->marksV, (uint32_t*)ws->marksR,
                                                       (uint32_t*)ws->pts, (BlobRecord*)ws->recs, rec_count, status);
        blob_contour_kernel<<<b3_blocks, kB3Threads, 0, stream>>>(g, (const uint32_t*)ws->planes, (const uint32_t*)ws->pts,
                                                                  (BlobRecord*)ws->recs, rec_count, (int*)ws->scratch, scratch_stride);
        if (e1) cudaEventRecord(e1, stream);
        BLOB_TRY(cudaGetLastError());
        unsigned hc[4];
        BLOB_TRY(cudaMemcpyAsync(hc, ws->counters, 16, cudaMemcpyDeviceToHost, stream));
        BLOB_TRY(cudaStreamSynchronize(stream));
        if (e0) { float t = 0; cudaEventElapsedTime(&t, e0, e1); *ms_out += t; }
        const bool overflow = hc[1] != 0 || hc[0] > g.rec_cap;
        if (!overflow)
        {
            ws->host_recs.resize(hc[0]);
            if (hc[0]) BLOB_TRY(cudaMemcpy(ws->host_recs.data(), ws->recs, sizeof(BlobRecord) * hc[0], cudaMemcpyDeviceToHost));
            break;
        }
        // A point region or the record list overflowed. A single frame is run again with four times the
        // space; a chunk of several frames is handed back to the caller, who runs its frames one by one
        // (growing the scratch for a whole chunk could ask for tens of GB because of one busy frame).
        if (n > 1) { if (e0) { cudaEventDestroy(e0); cudaEventDestroy(e1); } return 1; }
        if (attempt >= 6) { fprintf(stderr, "%s:%d in %s(): blob scratch still overflows after growing it 4096x. Sorry.\n", __FILE__, __LINE__, __func__); return -1; }
        ws->pts_cap *= 4; ws->rec_per_job *= 4;
    }
    if (e0) { cudaEventDestroy(e0); cudaEventDestroy(e1); }

    std::vector<BlobRecord>& R = ws->host_recs;
    std::sort(R.begin(), R.end(), [](const BlobRecord& a, const BlobRecord& b)
    {
        if (a.frame != b.frame) return a.frame < b.frame;
        if (a.thr != b.thr) return a.thr < b.thr;
        return a.seq > b.seq;
    });
    size_t i = 0;
    while (i < R.size())
    {
        size_t j = i;
        while (j < R.size() && R[j].frame == R[i].frame) j++;
        const int f = R[i].frame;
        counts_out[f] = group_frame(&R[i], (int)(j - i), xy_out + (size_t)f * 2 * max_points, max_points);
        i = j;
    }
    return 0;
}

}
