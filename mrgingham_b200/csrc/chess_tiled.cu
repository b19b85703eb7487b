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
#include <cuda_fp16.h>
#include <mutex>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>

#include "kernels.cuh"

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
constexpr int kTileThreads = 128;
constexpr uint32_t kFull = 0xffffffffu;

// ------------------------------------------------------------------------------------------------
// PTX helpers: mbarrier + TMA
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, 0x4000;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        :: "r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)) : "memory");
}

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

// The seven column offsets a pixel pair needs from one staged row, as [b0,0,b1,0] lanes.
struct RowSamples { uint32_t m5, m4, m2, c0, p2, p4, p5; };

// CLS = 0: pair at x = X (X = 0 mod 4); CLS = 2: pair at x = X+2. `row` points at the staged word
// holding bytes X-8..X-5 of this thread.
template<int CLS>
__device__ __forceinline__ RowSamples unpack_row(const uint8_t* row)
{
    RowSamples s;
    const uint32_t* wp = reinterpret_cast<const uint32_t*>(row);
    if (CLS == 0)
    {
        const uint32_t A = wp[0], B = wp[1], C = wp[2], D = wp[3];     // X-8, X-4, X, X+4
        s.m5 = __byte_perm(__byte_perm(A, B, 0x0043), 0, 0x4140);      // X-5 | X-4 straddles two words
        s.m4 = __byte_perm(B, 0, 0x4140);
        s.m2 = __byte_perm(B, 0, 0x4342);
        s.c0 = __byte_perm(C, 0, 0x4140);
        s.p2 = __byte_perm(C, 0, 0x4342);
        s.p4 = __byte_perm(D, 0, 0x4140);
        s.p5 = __byte_perm(D, 0, 0x4241);
    }
    else
    {
        const uint32_t B = wp[1], C = wp[2], D = wp[3], E = wp[4];     // X-4, X, X+4, X+8
        s.m5 = __byte_perm(B, 0, 0x4241);                              // X-3 | X-2
        s.m4 = __byte_perm(B, 0, 0x4342);
        s.m2 = __byte_perm(C, 0, 0x4140);
        s.c0 = __byte_perm(C, 0, 0x4342);
        s.p2 = __byte_perm(D, 0, 0x4140);
        s.p4 = __byte_perm(D, 0, 0x4342);
        s.p5 = __byte_perm(__byte_perm(D, E, 0x0043), 0, 0x4140);      // X+7 | X+8 straddles two words
    }
    return s;
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
    const int r0 = chess_from_rows(rows);
#pragma unroll
    for (int k = 0; k < 7; k++) rows[k] += 1;
    const int r1 = chess_from_rows(rows);
    const bool hit0 = r0 > kRespMin && x     >= kMargin && x     < w - kMargin;
    const bool hit1 = r1 > kRespMin && x + 1 >= kMargin && x + 1 < w - kMargin;
    const uint32_t b0 = __ballot_sync(kFull, hit0), b1 = __ballot_sync(kFull, hit1);
    if ((b0 | b1) == 0) return;
    const int lane = threadIdx.x & 31;
    uint32_t base = 0;
    if (lane == 0) base = atomicAdd(count, (uint32_t)(__popc(b0) + __popc(b1)));
    base = __shfl_sync(kFull, base, 0);
    const uint32_t below = (1u << lane) - 1;
    if (hit0)
    {
        const uint32_t i = base + __popc(b0 & below);
        if (i < (uint32_t)cap) out[i] = cand_pack(x, y, r0);
    }
    if (hit1)
    {
        const uint32_t i = base + __popc(b0) + __popc(b1 & below);
        if (i < (uint32_t)cap) out[i] = cand_pack(x + 1, y, r1);
    }
}

struct TileParams
{
    int nstrips, nsegs, seg_rows;   // work decomposition: item = (frame, segment, strip)
    int cap;
    int lookahead;                  // stages kept in flight ahead of the consumers (<= kStages-2)
};

template<int CLS, bool USE_TMA>
__device__ __forceinline__ void strip_walk(const CUtensorMap* tmap, const FrameSet& fs, const TileParams& tp,
                                           uint8_t* ring, uint64_t* full_bar, uint64_t* empty_bar, int* next_issue,
                                           cand_t* __restrict__ cand, uint32_t* __restrict__ counts)
{
    const int tid = threadIdx.x, lane = tid & 31, span = tid >> 6;
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

    const int X = xs + span * 128 + 4 * lane;            // this thread's word-aligned column
    const int x = X + CLS;                               // its pixel pair is (x, x+1)
    const int lane_off = span * 128 + 4 * lane + (kHalo - 8);   // byte offset of column X-8 in a staged row
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
        // no-TMA loader: all 128 threads copy the 10 x 272 bytes with bounds checks (zero fill)
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

    uint32_t Um2[kStageRows], U0[kStageRows], Up2[kStageRows];   // offsets -2, 0, +2 (needed at dy = +5 and -5)
    uint32_t U4m[kStageRows], U4p[kStageRows];                   // offsets -4, +4   (dy = +4 and -4)
    uint32_t U5m[kStageRows], U5p[kStageRows];                   // offsets -5, +5   (dy = +2, 0, -2)
#pragma unroll
    for (int j = 0; j < kStageRows; j++) { Um2[j] = U0[j] = Up2[j] = U4m[j] = U4p[j] = U5m[j] = U5p[j] = 0; }

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
            const RowSamples n = unpack_row<CLS>(stage + j * kRowBytes);
            // Rows outside [ys,ye) (window priming, segment tail) are computed like any other and
            // dropped below: cheaper than a test on the always-executed path.
            {
                // opposite ring samples (s_k, s_k+8), k = 0..7
                // the window slot of row R-k is (j + 11 - k) % 11; the new row R goes to slot j
                const uint32_t a0 = Up2[(j + 1) % 11],   b0 = n.m2;                    // (+2,-5) (-2,+5)
                const uint32_t a1 = U0[(j + 1) % 11],    b1 = n.c0;                    // ( 0,-5) ( 0,+5)
                const uint32_t a2 = Um2[(j + 1) % 11],   b2 = n.p2;                    // (-2,-5) (+2,+5)
                const uint32_t a3 = U4m[(j + 2) % 11],   b3 = U4p[(j + 10) % 11];      // (-4,-4) (+4,+4)
                const uint32_t a4 = U5m[(j + 4) % 11],   b4 = U5p[(j + 8) % 11];       // (-5,-2) (+5,+2)
                const uint32_t a5 = U5m[(j + 6) % 11],   b5 = U5p[(j + 6) % 11];       // (-5, 0) (+5, 0)
                const uint32_t a6 = U5m[(j + 8) % 11],   b6 = U5p[(j + 4) % 11];       // (-5,+2) (+5,-2)
                const uint32_t a7 = U4m[(j + 10) % 11],  b7 = U4p[(j + 2) % 11];       // (-4,+4) (+4,-4)
                const uint32_t p0 = a0 + b0, p1 = a1 + b1, p2 = a2 + b2, p3 = a3 + b3;
                const uint32_t p4 = a4 + b4, p5 = a5 + b5, p6 = a6 + b6, p7 = a7 + b7;
                // sum_response = sum_i |p_i - p_i+4|   (half2 lanes, exact: |values| <= 2040)
                const uint32_t w0 = h2sub(p0, p4), w1 = h2sub(p1, p5), w2 = h2sub(p2, p6), w3 = h2sub(p3, p7);
                const uint32_t s01 = h2absadd(w0, w1), s23 = h2absadd(w2, w3);
                const uint32_t sumr = h2absadd(s01, s23);
                // diff_response = sum_k |s_k - s_k+8|  (byte abs-diff on the zero-extended lanes).
                // Branch-free on purpose: a per-row early-out on `sumr` alone fired on ~40 % of the
                // warp-rows of board frames and its vote + branch cost more than it saved.
                // (three of the eight terms go through half2 instead of VABSDIFF4: the ALU pipe is
                // the busier one -- PRMT, VABSDIFF4, IADD3 -- so this evens out the two issue pipes)
                const uint32_t diff_h = h2absadd(h2absadd(h2sub(a0, b0), h2sub(a1, b1)), h2sub(a2, b2));
                const uint32_t diff = diff_h + __vabsdiffu4(a3, b3) +
                                      (__vabsdiffu4(a4, b4) + __vabsdiffu4(a5, b5)) + (__vabsdiffu4(a6, b6) + __vabsdiffu4(a7, b7));
                // lanes of (sumr - diff + 2048 + 0x77F0) reach 0x8000 iff sumr - diff >= 16, and
                // response = sumr - diff - |mean - local_mean| <= sumr - diff: only such rows can
                // hold a candidate; they are settled exactly after the block.
                const uint32_t t = sumr - diff + 0x7FF07FF0u;
                if (t & 0x80008000u) pending |= 1u << j;
            }
            Um2[j] = n.m2; U0[j] = n.c0; Up2[j] = n.p2; U4m[j] = n.m4; U4p[j] = n.p4; U5m[j] = n.m5; U5p[j] = n.p5;
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

template<bool USE_TMA>
__global__ void __launch_bounds__(kTileThreads)
chess_tiled_kernel(const __grid_constant__ CUtensorMap tmap, FrameSet fs, TileParams tp,
                   cand_t* __restrict__ cand, uint32_t* __restrict__ counts)
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
    // warps alternate between the two pixel-pair alignments of the same 128-pixel span
    if (((threadIdx.x >> 5) & 1) == 0) strip_walk<0, USE_TMA>(&tmap, fs, tp, ring, full_bar, empty_bar, &next_issue, cand, counts);
    else                               strip_walk<2, USE_TMA>(&tmap, fs, tp, ring, full_bar, empty_bar, &next_issue, cand, counts);
}

// ------------------------------------------------------------------------------------------------
// host: tensor map + launch
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn()
{
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, []
    {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    });
    return fn;
}

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

cudaError_t launch_chess_sparse_tiled(const FrameSet& fs, cand_t* cand, uint32_t* counts,
                                      int cand_capacity, cudaStream_t stream)
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
    int seg_rows = (int)((out_rows + nsegs - 1) / nsegs);
    seg_rows = ((seg_rows + kStageRows - 1) / kStageRows) * kStageRows;
    if (seg_rows < 40) seg_rows = 40;
    tp.seg_rows = seg_rows;
    tp.nsegs = (out_rows + seg_rows - 1) / seg_rows;
    const long long items = (long long)fs.nframes * tp.nstrips * tp.nsegs;
    if (items > 0x7fffffffLL) return cudaErrorInvalidValue;

    CUtensorMap map;
    if (make_tensor_map(&map, fs))
        chess_tiled_kernel<true><<<(unsigned)items, kTileThreads, 0, stream>>>(map, fs, tp, cand, counts);
    else
    {
        memset(&map, 0, sizeof(map));
        chess_tiled_kernel<false><<<(unsigned)items, kTileThreads, 0, stream>>>(map, fs, tp, cand, counts);
    }
    return cudaGetLastError();
}

}
