// K1, tiled variant: the production ChESS + candidate-emission kernel for sm_100a.
//
// Replaces, fused into one pass over the uint8 frame (1 byte/pixel of HBM traffic):
//   ChESS.c:62-105 (response), find_chessboard_corners.cc:506 (zeroed response), :527-529 (clamp),
//   and the r > 15 seed/member test of :159-171 -- only pixels with response > 15 are written out.
//
// Data movement. A CTA (4 warps) owns a 256-pixel-wide column strip of one frame and walks down
// it. Rows arrive through an 8-stage shared-memory ring filled by TMA (cp.async.bulk.tensor.3d,
// one 288-byte x 11-row box per stage: the strip plus a 16-byte halo each side), completion
// signalled on mbarriers; stages are handed back through a second set of mbarriers. Out-of-image
// rows/columns are zero-filled by TMA. Frames whose base/pitch do not meet TMA's 16-byte rules
// take the same kernel with a cooperative ld.global -> st.shared loader instead.
//
// Arithmetic. The bound is instruction issue, not HBM, so the work per pixel is
// minimised rather than the bytes:
//   * a thread owns TWO horizontally adjacent pixels, held as 2 x 16-bit lanes of a 32-bit
//     register ([b0,0,b1,0], one PRMT from the staged bytes). A non-negative value < 2048 in a
//     16-bit lane is simultaneously a valid integer AND a valid (subnormal/small) fp16 number with
//     the same bit pattern scaled by 2^-24, so integer adds (IADD3/VIADD), byte absolute
//     differences (VABSDIFF4) and half2 adds with free |x| / -x operand modifiers (HADD2) can be
//     mixed on the same registers, which spreads the work over both the ALU and the FMA pipe;
//   * warps are specialised by pixel-pair alignment (x = 0 or 2 mod 4) so that every ring sample
//     of a pair lies in ONE staged 32-bit word (two words for one of the seven column offsets);
//   * each thread walks down its column keeping the unpacked ring samples of the last 11 rows in
//     registers (rows y-5..y+5 are needed per output row, but only row y+5 is new), the row loop
//     is unrolled x11 so the window is addressed statically;
//   * exact early-out: response = sum - diff - |mean - local_mean| <= sum - diff, so the hot loop
//     forms only `sum` and `diff` (branch-free) and flags the rows where some lane has
//     sum - diff > 15 (~1 % of warp-rows on board frames); after each 11-row block the flagged rows
//     are recomputed exactly from the staged bytes by an out-of-line scalar routine that appends
//     the candidates. Whenever a response can exceed 15 it is computed exactly.
#include <cuda.h>
#include <algorithm>
#include <cuda_fp16.h>
#include <mutex>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>

#include "kernels.cuh"
#include "ptx.cuh"

namespace mrgb200
{

constexpr int kTileW      = 256;                 // output pixels per strip
constexpr int kHalo       = 16;                  // bytes staged left of the strip (and right): the ring needs 8,
                                                 // but TMA wants the box to start on a 16-byte boundary
constexpr int kRowBytes   = kTileW + 2*kHalo;    // 288
constexpr int kStageRows  = 11;                  // == unroll factor of the row loop == register-window length
constexpr int kStages     = 8;
constexpr int kLookahead  = 2;                   // stages requested ahead of the fastest warp. The ring is much
                                                 // deeper than that so the warp that issues never has to wait
                                                 // for a slower warp to hand a stage back.
constexpr int kStageBytes = 3200;                // 11*288 = 3168 rounded up to 128
constexpr int kTileThreads = 64;                // 2 warps x 32 lanes x 4 pixels = the 256-pixel strip

// ------------------------------------------------------------------------------------------------
// packed-lane helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t h2sub(uint32_t a, uint32_t b)
{
    __half2 r = __hsub2(*reinterpret_cast<__half2*>(&a), *reinterpret_cast<__half2*>(&b));
    return *reinterpret_cast<uint32_t*>(&r);
}
__device__ __forceinline__ uint32_t h2absadd(uint32_t a, uint32_t b)   // |a| + |b| per half lane
{
    __half2 r = __hadd2(__habs2(*reinterpret_cast<__half2*>(&a)), __habs2(*reinterpret_cast<__half2*>(&b)));
    return *reinterpret_cast<uint32_t*>(&r);
}

// The ten pixel pairs a thread needs from one staged row, as [b0,0,b1,0] lanes. A thread owns the
// four pixels X..X+3 of one staged 32-bit word: pair "0" = (X, X+1), pair "2" = (X+2, X+3).
// q_o is the pair (X+o, X+o+1). Pair 0 samples the ring at column offsets {-5,-4,-2,0,+2,+4,+5}
// -> q_-5, q_-4, q_-2, q_0, q_2, q_4, q_5; pair 2 at the same offsets from X+2
// -> q_-3, q_-2, q_0, q_2, q_4, q_6, q_7: four of the fourteen are shared.
struct RowPairs { uint32_t m5, m4, m3, m2, c0, p2, p4, p5, p6, p7;
                  uint32_t m1, p1, p3; };     // only filled in dense mode: the pairs local_mean needs

// `row` points at the staged word holding bytes X-8..X-5 of this thread.
template<bool DENSE>
__device__ __forceinline__ RowPairs unpack_row(const uint8_t* row)
{
    RowPairs q;
    const uint32_t* wp = reinterpret_cast<const uint32_t*>(row);
    const uint32_t A = wp[0], B = wp[1], C = wp[2], D = wp[3], E = wp[4];   // X-8, X-4, X, X+4, X+8
    q.m5 = __byte_perm(__byte_perm(A, B, 0x0043), 0, 0x4140);              // X-5 | X-4 straddles two words
    q.m4 = __byte_perm(B, 0, 0x4140);
    q.m3 = __byte_perm(B, 0, 0x4241);
    q.m2 = __byte_perm(B, 0, 0x4342);
    q.c0 = __byte_perm(C, 0, 0x4140);
    q.p2 = __byte_perm(C, 0, 0x4342);
    q.p4 = __byte_perm(D, 0, 0x4140);
    q.p5 = __byte_perm(D, 0, 0x4241);
    q.p6 = __byte_perm(D, 0, 0x4342);
    q.p7 = __byte_perm(__byte_perm(D, E, 0x0043), 0, 0x4140);              // X+7 | X+8 straddles two words
    if (DENSE)
    {
        q.m1 = __byte_perm(__byte_perm(B, C, 0x0043), 0, 0x4140);          // X-1 | X
        q.p1 = __byte_perm(C, 0, 0x4241);                                  // X+1 | X+2
        q.p3 = __byte_perm(__byte_perm(C, D, 0x0043), 0, 0x4140);          // X+3 | X+4
    }
    else q.m1 = q.p1 = q.p3 = 0;
    return q;
}

// Rare path (entered warp-uniformly, after the 11-row block that flagged it): some pixel of this
// warp's 64-pixel row segment may have response > 15. Recompute both pixels of every lane exactly,
// straight from the staged bytes (scalar, as ChESS.c:62-105 reads), and append the hits.
// rows[k] points at the staged byte of pixel x in row y + dy_k, dy = {-5,-4,-2,0,+2,+4,+5}.
__device__ __forceinline__ int chess_from_rows(const uint8_t* const* rows)
{
    const uint8_t *m5 = rows[0], *m4 = rows[1], *m2 = rows[2], *c = rows[3], *p2 = rows[4], *p4 = rows[5], *p5 = rows[6];
    const int s0 = m5[2],  s1 = m5[0],  s2  = m5[-2], s3  = m4[-4], s4  = m2[-5], s5  = c[-5], s6  = p2[-5], s7  = p4[-4];
    const int s8 = p5[-2], s9 = p5[0],  s10 = p5[2],  s11 = p4[4],  s12 = p2[5],  s13 = c[5],  s14 = m2[5],  s15 = m4[4];
    const int q0 = s0 + s8, q1 = s1 + s9, q2 = s2 + s10, q3 = s3 + s11, q4 = s4 + s12, q5 = s5 + s13, q6 = s6 + s14, q7 = s7 + s15;
    const int sum  = abs(q0 - q4) + abs(q1 - q5) + abs(q2 - q6) + abs(q3 - q7);
    const int diff = abs(s0 - s8) + abs(s1 - s9) + abs(s2 - s10) + abs(s3 - s11) + abs(s4 - s12) + abs(s5 - s13) + abs(s6 - s14) + abs(s7 - s15);
    const int mean = (q0 + q1 + q2 + q3) + (q4 + q5 + q6 + q7);
    const int local_mean = (c[-1] + c[0] + c[1]) * 16 / 3;
    return sum - diff - abs(mean - local_mean);
}

__device__ __noinline__ void emit_row(const uint8_t* cur_stage, const uint8_t* prev_stage, int j, int col /* byte of pixel x in a staged row */,
                                      int x, int y, int w, cand_t* __restrict__ out, uint32_t* __restrict__ count, int cap)
{
    // row y + dy is (5 - dy) rows behind the newest staged row (slot j of the current stage)
    const int behind[7] = { 10, 9, 7, 5, 3, 1, 0 };
    const uint8_t* rows[7];
#pragma unroll
    for (int k = 0; k < 7; k++)
    {
        const int r = j - behind[k];
        rows[k] = (r >= 0 ? cur_stage + r * kRowBytes : prev_stage + (r + kStageRows) * kRowBytes) + col;
    }
    const int lane = threadIdx.x & 31;
    const uint32_t below = (1u << lane) - 1;
    // the thread's four pixels x .. x+3, one ballot each
#pragma unroll 1
    for (int i = 0; i < 4; i++)
    {
        const int r = chess_from_rows(rows);
#pragma unroll
        for (int k = 0; k < 7; k++) rows[k] += 1;
        const bool hit = r > kRespMin && x + i >= kMargin && x + i < w - kMargin;
        const uint32_t b = __ballot_sync(kFull, hit);
        if (b == 0) continue;
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(count, (uint32_t)__popc(b));
        base = __shfl_sync(kFull, base, 0);
        if (hit)
        {
            const uint32_t idx = base + __popc(b & below);
            if (idx < (uint32_t)cap) out[idx] = cand_pack(x + i, y, r);
        }
    }
}

struct TileParams
{
    int nstrips, nsegs, seg_rows;   // work decomposition: item = (frame, segment, strip)
    int cap;
    int lookahead;                  // stages kept in flight ahead of the consumers (<= kStages-2)
};

// sum_i |p_i - p_i+4| and sum_k |a_k - b_k| for one pixel pair; returns the word whose lanes reach
// 0x8000 iff sum - diff >= 16 (a necessary condition for response > 15)
__device__ __forceinline__ uint32_t pair_test(uint32_t a0, uint32_t b0, uint32_t a1, uint32_t b1, uint32_t a2, uint32_t b2,
                                              uint32_t a3, uint32_t b3, uint32_t a4, uint32_t b4, uint32_t a5, uint32_t b5,
                                              uint32_t a6, uint32_t b6, uint32_t a7, uint32_t b7)
{
    const uint32_t p0 = a0 + b0, p1 = a1 + b1, p2 = a2 + b2, p3 = a3 + b3;
    const uint32_t p4 = a4 + b4, p5 = a5 + b5, p6 = a6 + b6, p7 = a7 + b7;
    // sum_response = sum_i |p_i - p_i+4|   (half2 lanes, exact: |values| <= 2040)
    const uint32_t sumr = h2absadd(h2absadd(h2sub(p0, p4), h2sub(p1, p5)), h2absadd(h2sub(p2, p6), h2sub(p3, p7)));
    // diff_response = sum_k |a_k - b_k|: three terms through half2, five through VABSDIFF4, which
    // evens out the ALU pipe (PRMT, VABSDIFF4, IADD3) and the FMA pipe (IMAD.IADD, HADD2/HFMA2)
    const uint32_t diff_h = h2absadd(h2absadd(h2sub(a0, b0), h2sub(a1, b1)), h2sub(a2, b2));
    const uint32_t diff = diff_h + __vabsdiffu4(a3, b3) +
                          (__vabsdiffu4(a4, b4) + __vabsdiffu4(a5, b5)) + (__vabsdiffu4(a6, b6) + __vabsdiffu4(a7, b7));
    // lanes of (sumr - diff + 2048 + 0x77F0) reach 0x8000 iff sumr - diff >= 16
    return sumr - diff + 0x7FF07FF0u;
}

// Dense mode: the exact response (ChESS.c:93-104) of one pixel pair from its sixteen ring-sample
// pairs (a_k, b_k) = (s_k, s_k+8) and the three centre pairs (x-1, x, x+1); two int16 in one word.
__device__ __forceinline__ uint32_t pair_response(uint32_t a0, uint32_t b0, uint32_t a1, uint32_t b1, uint32_t a2, uint32_t b2,
                                                  uint32_t a3, uint32_t b3, uint32_t a4, uint32_t b4, uint32_t a5, uint32_t b5,
                                                  uint32_t a6, uint32_t b6, uint32_t a7, uint32_t b7,
                                                  uint32_t cm1, uint32_t c0, uint32_t cp1)
{
    const uint32_t p0 = a0 + b0, p1 = a1 + b1, p2 = a2 + b2, p3 = a3 + b3;
    const uint32_t p4 = a4 + b4, p5 = a5 + b5, p6 = a6 + b6, p7 = a7 + b7;
    // sum_response, half2 lanes (exact: every value <= 2040)
    const uint32_t sumr = h2absadd(h2absadd(h2sub(p0, p4), h2sub(p1, p5)), h2absadd(h2sub(p2, p6), h2sub(p3, p7)));
    const uint32_t diff = (__vabsdiffu4(a0, b0) + __vabsdiffu4(a1, b1)) + (__vabsdiffu4(a2, b2) + __vabsdiffu4(a3, b3)) +
                          (__vabsdiffu4(a4, b4) + __vabsdiffu4(a5, b5)) + (__vabsdiffu4(a6, b6) + __vabsdiffu4(a7, b7));
    const uint32_t mean = ((p0 + p1) + (p2 + p3)) + ((p4 + p5) + (p6 + p7));     // <= 4080 per lane
    const uint32_t s3 = cm1 + c0 + cp1;                                          // <= 765 per lane
    // local_mean = s3*16/3 truncated = (s3 * 699056) >> 17 for s3 <= 765 (checked exhaustively)
    const int lm0 = (int)(((s3 & 0xFFFFu) * 699056u) >> 17), lm1 = (int)(((s3 >> 16) * 699056u) >> 17);
    const int r0 = (int)(sumr & 0xFFFFu) - (int)(diff & 0xFFFFu) - abs((int)(mean & 0xFFFFu) - lm0);
    const int r1 = (int)(sumr >> 16) - (int)(diff >> 16) - abs((int)(mean >> 16) - lm1);
    return ((uint32_t)r0 & 0xFFFFu) | ((uint32_t)r1 << 16);
}

template<bool USE_TMA, bool DENSE>
__device__ __forceinline__ void strip_walk(const CUtensorMap* tmap, const FrameSet& fs, const TileParams& tp,
                                           uint8_t* ring, uint64_t* full_bar, uint64_t* empty_bar, int* next_issue,
                                           cand_t* __restrict__ cand, uint32_t* __restrict__ counts,
                                           int16_t* __restrict__ response = nullptr, size_t resp_frame_stride = 0)
{
    const int tid = threadIdx.x, lane = tid & 31;
    const int item = blockIdx.x;
    const int strip = item % tp.nstrips;
    const int seg   = (item / tp.nstrips) % tp.nsegs;
    const int f     = item / (tp.nstrips * tp.nsegs);
    const int w = fs.w, h = fs.h;
    const int xs = strip * kTileW;                       // first output column of the strip
    const int ys = kMargin + seg * tp.seg_rows;          // first output row of the segment
    const int ye = min(ys + tp.seg_rows, h - kMargin);
    // Iteration `it`, step j stages row rbase + it*11 + j (the "+5" row of output row y = that - 5).
    // The first ten staged rows (ys-5 .. ys+4) only prime the register window.
    const int rbase = ys - 5;
    const int nit = (ye - ys + 10 + kStageRows - 1) / kStageRows;

    const int x = xs + 4 * tid;                          // this thread's pixels are x .. x+3 (one staged word)
    const int lane_off = 4 * tid + (kHalo - 8);          // byte offset of column x-8 in a staged row
    cand_t*   out   = cand + (size_t)f * tp.cap;
    uint32_t* count = counts + f;

    auto issue = [&](int it)
    {
        // one elected thread: refill the stage of iteration `it`
        const int s = it % kStages;
        if (it >= kStages) mbar_wait(&empty_bar[s], ((it / kStages) - 1) & 1);
        mbar_arrive_expect_tx(&full_bar[s], kRowBytes * kStageRows);
        tma_load_3d(ring + s * kStageBytes, tmap, (xs - kHalo) / 2, rbase + it * kStageRows, f, &full_bar[s]);
    };
    auto load_stage_cooperative = [&](int it)
    {
        // no-TMA loader: all threads copy the 11 x 288 bytes with bounds checks (zero fill)
        const int s = it % kStages;
        const uint8_t* img = fs.base + (size_t)f * fs.frame_stride;
        for (int i = tid; i < kStageRows * kRowBytes; i += kTileThreads)
        {
            const int r = i / kRowBytes, c = i % kRowBytes;
            const int gy = rbase + it * kStageRows + r, gx = xs - kHalo + c;
            uint8_t v = 0;
            if (gy >= 0 && gy < h && gx >= 0 && gx < w) v = img[(size_t)gy * fs.pitch + gx];
            ring[s * kStageBytes + i] = v;
        }
    };

    if (USE_TMA)
    {
        if (tid == 0)
            for (int it = 0; it < tp.lookahead && it < nit; it++) issue(it);
    }
    // Whichever warp reaches an iteration first requests the stage `lookahead` iterations ahead
    // (claimed through *next_issue), so no warp ever waits on a slower warp's producer duty.

    // Register window: pair q_o of the last 11 staged rows, for the offsets that are needed again
    // on later output rows. Ring row dy = +-5 uses q_{-2,0,2} (pair 0) / q_{0,2,4} (pair 2);
    // dy = +-4 uses q_{-4,4} / q_{-2,6}; dy = 0,+-2 uses q_{-5,5} / q_{-3,7}.
    uint32_t Wm5[kStageRows], Wm4[kStageRows], Wm3[kStageRows], Wm2[kStageRows], W0[kStageRows];
    uint32_t Wp2[kStageRows], Wp4[kStageRows], Wp5[kStageRows], Wp6[kStageRows], Wp7[kStageRows];
    uint32_t Wm1[kStageRows], Wp1[kStageRows], Wp3[kStageRows];     // dense mode: centre pairs of the row 5 steps back
#pragma unroll
    for (int j = 0; j < kStageRows; j++)
    {
        Wm5[j] = Wm4[j] = Wm3[j] = Wm2[j] = W0[j] = Wp2[j] = Wp4[j] = Wp5[j] = Wp6[j] = Wp7[j] = 0;
        Wm1[j] = Wp1[j] = Wp3[j] = 0;
    }
    int16_t* resp = DENSE ? response + (size_t)f * resp_frame_stride : nullptr;
    // 8-byte stores of four responses need aligned rows; border columns are written one by one
    const bool wide_ok = DENSE && !((uintptr_t)resp & 7) && !(w & 3);

#pragma unroll 1
    for (int it = 0; it < nit; it++)
    {
        const int s = it % kStages;
        if (USE_TMA)
        {
            if (lane == 0 && it + tp.lookahead < nit &&
                atomicCAS(next_issue, it + tp.lookahead, it + tp.lookahead + 1) == it + tp.lookahead)
                issue(it + tp.lookahead);
            mbar_wait(&full_bar[s], (it / kStages) & 1);
            __syncwarp();
        }
        else
        {
            __syncthreads();               // everyone is done with the stage being overwritten
            load_stage_cooperative(it);
            __syncthreads();
        }
        const uint8_t* stage = ring + s * kStageBytes + lane_off;
        const int ybase = rbase + it * kStageRows - 5;   // output row of step 0 (its +5 row is staged row 0)
        uint32_t pending = 0;                            // steps whose row may hold candidates (per lane until the block ends)

#pragma unroll
        for (int j = 0; j < kStageRows; j++)
        {
            const RowPairs n = unpack_row<DENSE>(stage + j * kRowBytes);
            // Rows outside [ys,ye) (window priming, segment tail) are computed like any other and
            // dropped below: cheaper than a test on the always-executed path.
            // The window slot of row R-k is (j + 11 - k) % 11; the new row R goes to slot j.
            constexpr int N = kStageRows;
            const int r10 = (j + 1) % N, r9 = (j + 2) % N, r7 = (j + 4) % N, r5 = (j + 6) % N, r3 = (j + 8) % N, r1 = (j + 10) % N;
            // opposite ring samples (s_k, s_k+8), k = 0..7:
            //   (+2,-5)(-2,+5)  (0,-5)(0,+5)  (-2,-5)(+2,+5)  (-4,-4)(+4,+4)
            //   (-5,-2)(+5,+2)  (-5,0)(+5,0)  (-5,+2)(+5,-2)  (-4,+4)(+4,-4)
            if (DENSE)
            {
                // all four responses of the word, exactly; only interior pixels are written (ChESS.c:62-63)
                const uint32_t ra = pair_response(Wp2[r10], n.m2,  W0[r10],  n.c0,  Wm2[r10], n.p2,  Wm4[r9],  Wp4[r1],
                                                  Wm5[r7],  Wp5[r3], Wm5[r5], Wp5[r5], Wm5[r3], Wp5[r7], Wm4[r1], Wp4[r9],
                                                  Wm1[r5], W0[r5], Wp1[r5]);
                const uint32_t rb = pair_response(Wp4[r10], n.c0,  Wp2[r10], n.p2,  W0[r10],  n.p4,  Wm2[r9],  Wp6[r1],
                                                  Wm3[r7],  Wp7[r3], Wm3[r5], Wp7[r5], Wm3[r3], Wp7[r7], Wm2[r1], Wp6[r9],
                                                  Wp1[r5], Wp2[r5], Wp3[r5]);
                const int y = ybase + j;
                if (y >= ys && y < ye && x < w - kMargin && x + 3 >= kMargin)
                {
                    int16_t* o = resp + (size_t)y * w + x;
                    if (wide_ok && x >= kMargin && x + 3 < w - kMargin)
                        *reinterpret_cast<uint2*>(o) = make_uint2(ra, rb);
                    else
                    {
                        const uint32_t v[2] = { ra, rb };
#pragma unroll
                        for (int i = 0; i < 4; i++)
                            if (x + i >= kMargin && x + i < w - kMargin) o[i] = (int16_t)(v[i >> 1] >> (16 * (i & 1)));
                    }
                }
            }
            else
            {
            const uint32_t t0 = pair_test(Wp2[r10], n.m2,  W0[r10],  n.c0,  Wm2[r10], n.p2,  Wm4[r9],  Wp4[r1],
                                          Wm5[r7],  Wp5[r3], Wm5[r5], Wp5[r5], Wm5[r3], Wp5[r7], Wm4[r1], Wp4[r9]);
            const uint32_t t2 = pair_test(Wp4[r10], n.c0,  Wp2[r10], n.p2,  W0[r10],  n.p4,  Wm2[r9],  Wp6[r1],
                                          Wm3[r7],  Wp7[r3], Wm3[r5], Wp7[r5], Wm3[r3], Wp7[r7], Wm2[r1], Wp6[r9]);
            // response <= sum - diff: only rows where some lane reaches the threshold can hold a
            // candidate; they are settled exactly after the block.
            if ((t0 | t2) & 0x80008000u) pending |= 1u << j;
            }
            Wm5[j] = n.m5; Wm4[j] = n.m4; Wm3[j] = n.m3; Wm2[j] = n.m2; W0[j] = n.c0;
            Wp2[j] = n.p2; Wp4[j] = n.p4; Wp5[j] = n.p5; Wp6[j] = n.p6; Wp7[j] = n.p7;
            if (DENSE) { Wm1[j] = n.m1; Wp1[j] = n.p1; Wp3[j] = n.p3; }
        }

        pending = __reduce_or_sync(kFull, pending);
        while (pending)
        {
            const int j = __ffs(pending) - 1;
            pending &= pending - 1;
            const int y = ybase + j;
            if (y >= ys && y < ye)
                emit_row(ring + s * kStageBytes, ring + ((it + kStages - 1) % kStages) * kStageBytes, j,
                         kHalo + (x - xs), x, y, w, out, count, tp.cap);
        }

        if (USE_TMA && it >= 1)
        {
            // hand back the PREVIOUS stage: the rare path above still reads centre rows from it
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[(it - 1) % kStages]);
        }
    }
}

template<bool USE_TMA, bool DENSE>
__global__ void __launch_bounds__(kTileThreads, DENSE ? 4 : 8)
chess_tiled_kernel(const __grid_constant__ CUtensorMap tmap, FrameSet fs, TileParams tp,
                   cand_t* __restrict__ cand, uint32_t* __restrict__ counts,
                   int16_t* __restrict__ response, size_t resp_frame_stride)
{
    __shared__ __align__(128) uint8_t ring[kStages * kStageBytes];
    __shared__ __align__(8) uint64_t full_bar[kStages], empty_bar[kStages];
    __shared__ int next_issue;
    if (USE_TMA)
    {
        if (threadIdx.x == 0)
        {
            next_issue = tp.lookahead;   // the prologue below requests iterations 0 .. lookahead-1
            for (int s = 0; s < kStages; s++) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], kTileThreads / 32); }
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
    }
    strip_walk<USE_TMA, DENSE>(&tmap, fs, tp, ring, full_bar, empty_bar, &next_issue, cand, counts, response, resp_frame_stride);
}

// ------------------------------------------------------------------------------------------------
// host: tensor map + launch
// ------------------------------------------------------------------------------------------------
static bool make_tensor_map(CUtensorMap* map, const FrameSet& fs)
{
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) return false;
    if (((uintptr_t)fs.base & 15) || (fs.pitch & 15) || (fs.w & 1)) return false;
    size_t fstride = fs.frame_stride;
    if (fs.nframes == 1) fstride = ((size_t)fs.pitch * fs.h + 15) & ~(size_t)15;
    if (fstride & 15) return false;
    const cuuint64_t dims[3]    = { (cuuint64_t)(fs.w / 2), (cuuint64_t)fs.h, (cuuint64_t)fs.nframes };
    const cuuint64_t strides[2] = { (cuuint64_t)fs.pitch, (cuuint64_t)fstride };
    const cuuint32_t box[3]     = { kRowBytes / 2, kStageRows, 1 };
    const cuuint32_t estr[3]    = { 1, 1, 1 };
    const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT16, 3, (void*)fs.base, dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

static cudaError_t launch_tiled(const FrameSet& fs, cand_t* cand, uint32_t* counts, int cand_capacity,
                                int16_t* response, size_t resp_frame_stride, cudaStream_t stream)
{
    if (fs.w <= 2*kMargin || fs.h <= 2*kMargin || fs.nframes <= 0) return cudaSuccess;
    TileParams tp;
    tp.cap = cand_capacity;
    tp.lookahead = kLookahead;
    if (const char* e = getenv("MRG_B200_K1_LOOKAHEAD")) { int v = atoi(e); if (v >= 1 && v <= kStages - 2) tp.lookahead = v; }
    tp.nstrips = (fs.w - kMargin + kTileW - 1) / kTileW;       // strips start at x = 0
    const int out_rows = fs.h - 2*kMargin;
    // enough work items to fill the chip a few times over, but segments no shorter than 40 rows
    const long long want_items = 148LL * 5 * 4;
    long long nsegs = (want_items + (long long)fs.nframes * tp.nstrips - 1) / ((long long)fs.nframes * tp.nstrips);
    // ~275-row segments balance better than one segment per strip (measured on the cascade kernel)
    nsegs = std::max(nsegs, (long long)((out_rows + 274) / 275));
    int seg_rows = (int)((out_rows + nsegs - 1) / nsegs);
    seg_rows = ((seg_rows + kStageRows - 1) / kStageRows) * kStageRows;
    if (seg_rows < 40) seg_rows = 40;
    tp.seg_rows = seg_rows;
    tp.nsegs = (out_rows + seg_rows - 1) / seg_rows;
    const long long items = (long long)fs.nframes * tp.nstrips * tp.nsegs;
    if (items > 0x7fffffffLL) return cudaErrorInvalidValue;

    CUtensorMap map;
    const bool tma = make_tensor_map(&map, fs);
    if (!tma) memset(&map, 0, sizeof(map));
    const unsigned grid = (unsigned)items;
    if (response)
    {
        if (tma) chess_tiled_kernel<true,  true><<<grid, kTileThreads, 0, stream>>>(map, fs, tp, nullptr, nullptr, response, resp_frame_stride);
        else     chess_tiled_kernel<false, true><<<grid, kTileThreads, 0, stream>>>(map, fs, tp, nullptr, nullptr, response, resp_frame_stride);
    }
    else
    {
        if (tma) chess_tiled_kernel<true,  false><<<grid, kTileThreads, 0, stream>>>(map, fs, tp, cand, counts, nullptr, 0);
        else     chess_tiled_kernel<false, false><<<grid, kTileThreads, 0, stream>>>(map, fs, tp, cand, counts, nullptr, 0);
    }
    return cudaGetLastError();
}

cudaError_t launch_chess_sparse_tiled(const FrameSet& fs, cand_t* cand, uint32_t* counts,
                                      int cand_capacity, cudaStream_t stream)
{
    return launch_tiled(fs, cand, counts, cand_capacity, nullptr, 0, stream);
}

// Dense int16 response (mrgingham_ChESS_response_5, ChESS.c:55-106) with the tiled kernel's machinery:
// every interior pixel's exact response, 1 byte read + 2 bytes written per pixel.
cudaError_t launch_chess_dense_tiled(const FrameSet& fs, int16_t* response, size_t response_frame_stride_elems,
                                     cudaStream_t stream)
{
    return launch_tiled(fs, nullptr, nullptr, 0, response, response_frame_stride_elems, stream);
}

}
