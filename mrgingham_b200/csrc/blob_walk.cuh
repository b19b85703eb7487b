// Border following on a bit plane as a permutation of border STATES, shared by the CUDA kernels of
// blobs.cu and by the host-side check tests/test_blob_walk_host.py compiles (plain C++ there).
//
// What cv::findContours(RETR_LIST, CHAIN_APPROX_NONE) does (Suzuki-Abe with a zero frame, 8-connected
// foreground; the test oracle restates it with its pixel marks) is equivalent to this mark-free
// formulation, which is what lets every border be handled by its own lane:
//
//   * directions d = 0..7: E, NE, N, NW, W, SW, S, SE (counter-clockwise on the screen, y down).
//   * a STATE is (pixel P, direction k) with N_k(P) foreground, N_{k-1}(P) background, and the maximal
//     run of background neighbours that ends at k-1 (going clockwise from k-1) holding an axial
//     (4-adjacent) neighbour. The border following visits P once per such run and leaves towards N_k(P).
//   * successor of (P,k): Q = N_k(P); search counter-clockwise from the direction after the one pointing
//     back to P; the first foreground neighbour is the new k. Predecessor: first foreground neighbour
//     clockwise from k-1 is where the trace came from. Both are bijections: the states fall into cycles,
//     one cycle per border (outer border of an 8-connected component, or border of a hole).
//   * the raster scan discovers a cycle at the smallest "discovery position" of its states: a state whose
//     run holds W is met at the 0->1 transition at P (position y*w + x); a state whose run holds E is met
//     at the 1->0 transition right of P (position y*w + x + 1, if x + 1 < w). The cycle is an outer border
//     if that minimum is of the first kind (it is then the component's first pixel), a hole border
//     otherwise (the pixel left of the hole's first pixel); the contour's points are the pixels of the
//     cycle's states, in cycle order, from that state on. Contours are listed in discovery order.
//
// tests/test_blob_walk_host.py checks exactly this (point lists, order, area sums) against the oracle's
// Suzuki-Abe restatement on random binary images.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define BW_HD __host__ __device__ __forceinline__
#else
#define BW_HD inline
#endif

namespace mrgb200
{
namespace blobwalk
{
BW_HD int ffs32(uint32_t v)
{
#if defined(__CUDA_ARCH__)
    return __ffs((int)v);
#else
    return __builtin_ffs((int)v);
#endif
}
BW_HD int clz32(uint32_t v)
{
#if defined(__CUDA_ARCH__)
    return __clz((int)v);
#else
    return v ? __builtin_clz(v) : 32;
#endif
}
// dx, dy of direction d, from two packed tables of (value + 1) in 2-bit fields
BW_HD int dir_dx(int d) { return (int)((0x901Au >> (2 * d)) & 3u) - 1; }
BW_HD int dir_dy(int d) { return (int)((0xA901u >> (2 * d)) & 3u) - 1; }

// One lane's view of a bit plane (bit b of word wd of row y = pixel (32*wd + b, y) is foreground; pixels
// outside the w x h image are background). The lane keeps the 3 x 3 words around its position in
// registers: while a border is followed most steps need no load at all, a vertical step needs one round
// of three independent loads.
// A bit plane as the walkers see it: B[y * wpr + wd], bit b = pixel (32*wd + b, y) is foreground. The storage
// carries a frame of background around the image, so that B may be read for y in [-1, h] and wd in [-1, words]
// (words = ceil(w / 32)) without any bounds test: B points at the image's first word, one row and one word into
// the storage, and bits at x >= w are 0.
struct PlaneRef
{
    const uint32_t* B;
    int w, h, wpr;
};
// storage words per row / rows of a plane holding a w x h image
BW_HD int plane_wpr(int w)  { return ((w + 31) / 32 + 2 + 3) & ~3; }
BW_HD int plane_rows(int h) { return h + 2; }
BW_HD int plane_origin(int w) { return plane_wpr(w) + 1; }           // word offset of pixel (0,0) in the storage

// One lane's view of a bit plane: the 3 x 3 words around its position, in registers. While a border is followed
// most steps need no load at all, a vertical step needs one round of three independent loads.
struct BitWindow
{
    int cy, cwd;
    uint32_t b[3][3];

    BW_HD void init() { cy = -0x40000000; cwd = -2; }
    BW_HD void load_row(const PlaneRef& P, int slot, int y)
    {
        const uint32_t* r = P.B + (y * P.wpr + cwd);
        b[slot][0] = r[-1]; b[slot][1] = r[0]; b[slot][2] = r[1];
    }
    BW_HD void seek(const PlaneRef& P, int x, int y)
    {
        const int wd = x >> 5;
        if (wd == cwd && y == cy) return;
        if (wd == cwd && y == cy + 1)
        {
#pragma unroll
            for (int c = 0; c < 3; c++) { b[0][c] = b[1][c]; b[1][c] = b[2][c]; }
            cy = y; load_row(P, 2, y + 1);
        }
        else if (wd == cwd && y == cy - 1)
        {
#pragma unroll
            for (int c = 0; c < 3; c++) { b[2][c] = b[1][c]; b[1][c] = b[0][c]; }
            cy = y; load_row(P, 0, y - 1);
        }
        else
        {
            cwd = wd; cy = y;
            load_row(P, 0, y - 1); load_row(P, 1, y); load_row(P, 2, y + 1);
        }
    }
    // bits (x-1, x, x+1) of a cached row as bits 0..2
    BW_HD uint32_t row3(int slot, int bpos) const
    {
        const unsigned long long w64 = ((unsigned long long)b[slot][1] << 32) | b[slot][0];
        uint32_t v = (uint32_t)(w64 >> (31 + bpos)) & 7u;
        if (bpos == 31) v |= (b[slot][2] & 1u) << 2;
        return v;
    }
    // bit d = the neighbour of (x, y) in direction d is foreground. x in [0, w), y in [0, h).
    BW_HD uint32_t nbr8(const PlaneRef& P, int x, int y)
    {
        seek(P, x, y);
        const int bpos = x & 31;
        const uint32_t up = row3(0, bpos), mid = row3(1, bpos), dn = row3(2, bpos);
        return ((mid >> 2) & 1u) | (((up >> 2) & 1u) << 1) | (((up >> 1) & 1u) << 2) | ((up & 1u) << 3) |
               ((mid & 1u) << 4) | ((dn & 1u) << 5) | (((dn >> 1) & 1u) << 6) | (((dn >> 2) & 1u) << 7);
    }
};

// (x,y,k) -> its successor state. *disc = the successor's discovery position, or -1 if the raster scan
// cannot meet it (its run of background neighbours holds neither W nor an E inside the image).
BW_HD void step_fwd(const PlaneRef& P, BitWindow& W, int& x, int& y, int& k, int* disc)
{
    const int qx = x + dir_dx(k), qy = y + dir_dy(k);
    const uint32_t m = W.nbr8(P, qx, qy);
    const int s0 = (k + 5) & 7;                                    // the direction after the one pointing back
    const uint32_t rot = ((m | (m << 8)) >> s0) & 0xFFu;           // bit j = direction s0 + j; bit 7 (back) is set
    const int j = ffs32(rot) - 1;                                  // background neighbours passed: directions s0 .. s0+j-1
    x = qx; y = qy; k = (s0 + j) & 7;
    int d = -1;
    if (((4 - s0) & 7) < j) d = qy * P.w + qx;
    else if (((0 - s0) & 7) < j && qx + 1 < P.w) d = qy * P.w + qx + 1;
    *disc = d;
}

// *disc = discovery position of the state (x,y,k) itself (or -1), then (x,y,k) -> its predecessor state.
BW_HD void step_bwd(const PlaneRef& P, BitWindow& W, int& x, int& y, int& k, int* disc)
{
    const uint32_t m = W.nbr8(P, x, y);
    const uint32_t r = ((m | (m << 8)) >> k) & 0xFFu;              // bit i = direction k + i; bit 0 (k itself) is set
    const int hb = 31 - clz32(r);                                  // first foreground neighbour clockwise from k-1
    // run of the state = directions k+hb+1 .. k+7
    int d = -1;
    if (((4 - k) & 7) > hb) d = y * P.w + x;
    else if (((0 - k) & 7) > hb && x + 1 < P.w) d = y * P.w + x + 1;
    *disc = d;
    const int a = (k + hb) & 7;
    x += dir_dx(a); y += dir_dy(a); k = (a + 4) & 7;
}

// ---- borders in SEGMENTS --------------------------------------------------------------------------------------
// A long border would keep the lane that follows it busy long after every other lane has finished, so borders are
// cut at SEGMENT STARTS, states that can be recognised locally both by a scan of the plane and by a walker that
// arrives at them:
//   * the candidate starts below (first pixel of a component / of a hole),
//   * row cuts: a state whose run holds W or E, at a pixel with y % 256 == 0,
//   * column cuts: a state whose run holds N or S, at a pixel with x % 256 == 0.
// A border cannot run further than that down or up a side, or along a top or a bottom, without meeting one, so
// segments stay short. Each
// segment start is walked FORWARD to the next one by its own lane (length, area sum, smallest discovery position of
// its states); a candidate then follows the chain of segments from its own -- one hash look-up per segment instead
// of one step per state -- until it is back at itself (it is where the scan discovers the border) or meets a
// segment holding a state the scan reaches earlier (it is not).
constexpr int kCutMask = 255;

// the state of pixel (x,y) whose run of background neighbours holds direction d0 (which must be background):
// *k = where the trace leaves; false if the pixel has no foreground neighbour at all
BW_HD bool state_after(const PlaneRef& P, BitWindow& W, int x, int y, int d0, int* k)
{
    const uint32_t m = W.nbr8(P, x, y);
    const int s0 = (d0 + 1) & 7;
    const uint32_t r = ((m | (m << 8)) >> s0) & 0x7Fu;             // directions d0+1 .. d0+7
    if (!r) return false;
    *k = (s0 + ffs32(r) - 1) & 7;
    return true;
}

// discovery position of the state (x,y,k) (or -1), whether it is a segment start, and, if it is a candidate start
// (first pixel of a component: its run holds W, NW, N, NE; pixel left of a hole's first: it leaves towards NE with E in
// its run), the position at which the raster scan would discover a border there (else -1)
BW_HD void classify_state(const PlaneRef& P, BitWindow& W, int x, int y, int k, int* disc, bool* is_start, int* cand_pos)
{
    const uint32_t m = W.nbr8(P, x, y);
    const uint32_t r = ((m | (m << 8)) >> k) & 0xFFu;              // bit i = direction k + i; bit 0 (k itself) is set
    const int hb = 31 - clz32(r);                                  // run of the state = directions k+hb+1 .. k+7
    const bool hasW = ((4 - k) & 7) > hb, hasN = ((2 - k) & 7) > hb, hasE = ((0 - k) & 7) > hb, hasS = ((6 - k) & 7) > hb;
    const bool hasNW = ((3 - k) & 7) > hb, hasNE = ((1 - k) & 7) > hb;
    int d = -1;
    if (hasW) d = y * P.w + x;
    else if (hasE && x + 1 < P.w) d = y * P.w + x + 1;
    *disc = d;
    const bool outer = hasW && hasNW && hasN && hasNE, hole = k == 1 && hasE && x + 1 < P.w;
    *cand_pos = outer ? y * P.w + x : hole ? y * P.w + x + 1 : -1;
    *is_start = outer || (k == 1 && hasE) || ((hasW || hasE) && (y & kCutMask) == 0) || ((hasN || hasS) && (x & kCutMask) == 0);
}

// (x,y,k) -> its successor, with the successor's discovery position and whether it is a segment start
BW_HD void step_fwd_ex(const PlaneRef& P, BitWindow& W, int& x, int& y, int& k, int* disc, bool* is_start)
{
    const int qx = x + dir_dx(k), qy = y + dir_dy(k);
    const uint32_t m = W.nbr8(P, qx, qy);
    const int s0 = (k + 5) & 7;
    const uint32_t rot = ((m | (m << 8)) >> s0) & 0xFFu;
    const int j = ffs32(rot) - 1;                                  // run of the new state = directions s0 .. s0+j-1
    x = qx; y = qy; k = (s0 + j) & 7;
    const bool hasW = ((4 - s0) & 7) < j, hasN = ((2 - s0) & 7) < j, hasE = ((0 - s0) & 7) < j, hasS = ((6 - s0) & 7) < j;
    const bool hasNW = ((3 - s0) & 7) < j, hasNE = ((1 - s0) & 7) < j;
    int d = -1;
    if (hasW) d = qy * P.w + qx;
    else if (hasE && qx + 1 < P.w) d = qy * P.w + qx + 1;
    *disc = d;
    *is_start = (hasW && hasNW && hasN && hasNE) || (k == 1 && hasE) || ((hasW || hasE) && (qy & kCutMask) == 0) || ((hasN || hasS) && (qx & kCutMask) == 0);
}

// key of a state within a chunk: job << 33 | y << 18 | x << 3 | k
BW_HD unsigned long long state_key(int job, int x, int y, int k)
{
    return (unsigned long long)job << 33 | (unsigned long long)y << 18 | (unsigned long long)x << 3 | (unsigned long long)k;
}

// cut masks of one plane word (cur / left / right = words wd, wd-1, wd+1 of row y; up / dn = word wd of rows y-1, y+1):
// foreground pixels with background to the W / to the E on rows y % 256 == 0, and with background to the N / to the S
// at x % 256 == 0. The state meant is the one whose run holds that neighbour (state_after with d0 = 4 / 0 / 2 / 6).
BW_HD void cut_masks(uint32_t cur, uint32_t left, uint32_t right, uint32_t up, uint32_t dn, int wd, int y,
                     uint32_t* cutW, uint32_t* cutE, uint32_t* cutN, uint32_t* cutS)
{
    const uint32_t Wn = (cur << 1) | (left >> 31), En = (cur >> 1) | (right << 31);
    const bool row = (y & kCutMask) == 0, col = ((wd * 32) & kCutMask) == 0;
    *cutW = row ? cur & ~Wn : 0u;
    *cutE = row ? cur & ~En : 0u;
    *cutN = col ? cur & ~up & 1u : 0u;
    *cutS = col ? cur & ~dn & 1u : 0u;
}

// Candidate starts, the only places a cycle's smallest discovery position can be:
//   kind 0: foreground pixel (x,y) whose W, NW, N, NE neighbours are background (first pixel of a component);
//           *k = where the trace leaves it, false for an isolated pixel (a one-point contour: area 0, never kept)
//   kind 1: background pixel (x,y) whose W and N neighbours are foreground (first pixel of a hole); the state is
//           (x-1, y, NE)
BW_HD bool outer_start(const PlaneRef& P, BitWindow& W, int x, int y, int* k)
{
    const uint32_t m = W.nbr8(P, x, y);
    const uint32_t r4 = ((m | (m << 8)) >> 5) & 0xFu;              // directions SW, S, SE, E
    if (!r4) return false;
    *k = (5 + ffs32(r4) - 1) & 7;
    return true;
}

// Candidate masks of one plane word: cur / left = words wd, wd-1 of row y; up / upleft / upright = words wd, wd-1,
// wd+1 of row y-1 (0 outside the plane). Bits beyond the image width come out 0 in both masks.
BW_HD void candidate_masks(uint32_t cur, uint32_t left, uint32_t up, uint32_t upleft, uint32_t upright, uint32_t* outer, uint32_t* hole)
{
    const uint32_t Wn = (cur << 1) | (left >> 31), NW = (up << 1) | (upleft >> 31), NE = (up >> 1) | (upright << 31);
    *outer = cur & ~Wn & ~NW & ~up & ~NE;
    *hole  = ~cur & Wn & up;
}

// Host-side (and reference) form of the verification walk: is (x,y,k), discovered at `pos`, the state its
// cycle is discovered at? Two walkers leave it in opposite directions and stop as soon as either meets a
// state the scan reaches earlier; if they meet each other instead the cycle is this candidate's. *n = states
// of the cycle, *a00 = sum over its directed edges P->Q of (Px*Qy - Qx*Py) (twice the signed area).
BW_HD bool verify_start(const PlaneRef& P, BitWindow& F, BitWindow& Bk, int x, int y, int k, int pos, int* n, long long* a00, long long max_steps)
{
    int fx = x, fy = y, fk = k, bx = x, by = y, bk = k, cnt = 0, disc;
    long long a = 0;
    for (long long it = 0; it < max_steps; it++)
    {
        { const int px = fx, py = fy; step_fwd(P, F, fx, fy, fk, &disc); a += (long long)(px * fy - fx * py); cnt++; }
        // (the backward walker's latest state has not been looked at yet: test before the meeting check)
        if (disc >= 0 && disc < pos) return false;
        if (fx == bx && fy == by && fk == bk) { *n = cnt; *a00 = a; return true; }
        { const int qx = bx, qy = by; step_bwd(P, Bk, bx, by, bk, &disc); a += (long long)(bx * qy - qx * by); cnt++; }
        if (disc >= 0 && disc < pos) return false;
        if (fx == bx && fy == by && fk == bk) { *n = cnt; *a00 = a; return true; }
    }
    return false;
}
}   // namespace blobwalk
}   // namespace mrgb200
