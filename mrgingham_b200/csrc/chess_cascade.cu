// K1, cascade variant: the production ChESS + candidate-emission kernel for sm_100a.
//
// Replaces, fused into one pass over the uint8 frame (1 byte/pixel of HBM traffic):
//   ChESS.c:62-105 (response), find_chessboard_corners.cc:506 (zeroed response), :527-529 (clamp),
//   and the r > 15 seed/member test of :159-171 -- only pixels with response > 15 are written out.
//
// The kernel is bound by instruction issue, not by HBM, so it is organised as a cascade of EXACT
// necessary conditions for `response > 15`, each cheaper and applied to more pixels than the next:
//
//   With a,b,c,d = s_i, s_i+4, s_i+8, s_i+12 (ChESS.c:93-102), term_i = |a+c-b-d| - |a-c| - |b-d|
//   and response = sum_i term_i - |mean - local_mean| <= sum_i term_i. For every i
//       term_i = 2*max( min(a,c) - max(b,d), min(b,d) - max(a,c) )                       (identity)
//   hence  term_i <= 2*|p - q|  for ANY p in {a,c}, q in {b,d},                           (L1)
//   and    term_i <= |a-b| + |c-d| - |a-c| - |b-d|  (triangle inequality on |a+c-b-d|).   (L2)
//
//   L1 (every pixel; 4 pixels per instruction in byte lanes): one chord |p-q| per i, chosen so
//      that three of the eight samples are aligned 32-bit words of the staged rows and the other
//      five come from two PRMT'd words per row. response > 15 needs sum_i |p_i-q_i| >= 8. The four
//      VABSDIFF4 results are added as packed bytes; a lane can only wrap if some chord is large, and
//      then the sum of all the word's chord bytes (IDP.4A, FMA pipe) is >= 32 and flags the word, so
//      the packed test never misses. ~5 % of the pixels of a board frame (edge neighbourhoods) pass.
//   L2 (8-pixel row cells flagged by L1, compacted so that all 32 lanes work): all sixteen
//      VABSDIFF4 chords/diameters in byte lanes, summed exactly per pixel with IDP.4A dot products
//      (FMA pipe); needs sum >= 16. ~0.05 % of the pixels pass.
//   L3 (those pixels, compacted again): the exact scalar response as ChESS.c:62-105 computes it;
//      pixels with response > 15 inside [7,w-7) x [7,h-7) are appended to the frame's candidate list.
//
// Whatever L1/L2 let through is only ever MORE than needed; the candidate set equals the
// reference's on any input (tests/test_gpu_parity.py compares it with the oracle on noise, textures
// and dense checkers, where nearly everything reaches L2).
//
// Data movement: a CTA of NW warps owns a (256*NW)-pixel-wide column strip of one frame segment
// and walks down it; rows arrive through a shared-memory ring filled by TMA
// (cp.async.bulk.tensor.3d, one (256*NW+32)-byte x 11-row box per stage), completion on mbarriers,
// stages handed back through a second set of mbarriers. A lane owns 8 adjacent pixels (two 32-bit
// words) and keeps the PRMT'd words of the last 11 rows in registers (row loop unrolled x11).
// L2 reads its cells from the two staged blocks still resident in shared memory (a stage is handed
// back two blocks late for that); L3 re-reads its few pixels from global memory.
#include <cuda.h>
#include <algorithm>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>

#include "kernels.cuh"
#include "ptx.cuh"

namespace mrgb200
{
namespace
{
constexpr int kLanePx   = 8;                   // pixels per lane
constexpr int kWarpPx   = 32 * kLanePx;        // 256
constexpr int kCHalo    = 16;                  // staged bytes left/right of the strip: the ring needs 5 + the lanes' aligned 8-byte cells, and a
                                               // TMA box must start on a 16-byte boundary of global memory (anything else is an illegal instruction)
constexpr int kBlkRows  = 11;                  // rows per TMA stage == unroll of the row loop == register window
constexpr int kL2QCarry = 512;                 // per-warp queue of flagged 8-pixel row cells: >= 31 carried over + 352 of one block, a power of two (slot = position & 511) ...
constexpr int kL2QBlock = 352;                 // ... or one block's worth when nothing is carried (the queue then restarts at 0 every block)
constexpr int kL3QCap   = 288;                 // per-warp queue of pixels for the exact test (>= 31 waiting + 256 of one L2 batch)

template<int NW> struct Geo
{
    static constexpr int kStripW   = kWarpPx * NW;
    // A staged row is the strip plus its halos, rounded up to an ODD number of 16-byte units (TMA boxes come in multiples
    // of 16 bytes): the lanes of an L2 batch read the same columns of up to 11 consecutive rows (cells along a vertical
    // edge), and with 800-byte rows every fourth row starts in the same bank -- 3.1 wavefronts per 64-bit load measured
    // where 1 would do; with 816 bytes it is every eighth row.
    static constexpr int kRowBytes = ((kStripW + 2 * kCHalo) / 16 % 2 == 0) ? kStripW + 2 * kCHalo + 16 : kStripW + 2 * kCHalo;
    static constexpr int kStageBytes = ((kRowBytes * kBlkRows + 127) / 128) * 128;
};

struct CascadeParams
{
    int nstrips, nsegs, seg_rows;   // work decomposition: item = (frame, segment, strip)
    int cap;
    int stages;                     // depth of the shared-memory ring (TMA stages of 11 rows)
};

// head/tail count entries since the CTA started (a CTA never queues anywhere near 2^32 of them); the slot of
// entry p is p % cap. Only the carried L2 queue's capacity is a power of two (512 instead of the 384 it needs: 5 CTAs
// still fit a SM, and the slot arithmetic of the flagged-cell loop gets shorter: +0.5-1 %); the others are as small as
// they can be, because every KB of shared memory per warp decides how many CTAs fit a SM.
// CARRY: flagged cells that do not fill a batch of 32 wait for the next block's (three staged blocks are then held,
// four stages needed); !CARRY: every block settles its own cells, the last batch partly empty (two held, three stages).
template<bool CARRY> struct WarpQueues
{
    static constexpr int kL2QCap = CARRY ? kL2QCarry : kL2QBlock;
    uint32_t l3q[kL3QCap];
    uint16_t l2q[kL2QCap];
    uint32_t l2_tail, l3_tail;
};

// mbarrier wait as a plain C loop around try_wait, executed by whole warps: no branch hidden from
// the compiler, so no lane is ever left diverged from its warp across the unrolled row loop
__device__ __forceinline__ void mbar_wait_warp(uint64_t* bar, uint32_t parity)
{
    uint32_t ok;
    do
    {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, 0x4000;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!ok);
}

__device__ __forceinline__ uint32_t vabs4(uint32_t a, uint32_t b) { return __vabsdiffu4(a, b); }

// u0 + u1 + u2 + u3 + 0x78787878 in packed bytes: two IADD3. (Moving these sums to the FMA pipe as IMADs by an
// opaque 1, and register targets of 88 / 80 for 6 / 7 CTAs per SM, were measured and lost 0-4 %:
// profiles/r02_k1_ab_knobs.txt.)
__device__ __forceinline__ uint32_t packed_sum4(uint32_t u0, uint32_t u1, uint32_t u2, uint32_t u3)
{
    return u0 + u1 + u2 + u3 + 0x78787878u;
}

// exact response of the pixel at c (ChESS.c:62-105), read straight from global memory
__device__ __forceinline__ int chess_exact(const uint8_t* __restrict__ c, int pitch)
{
    const int p2 = 2*pitch, p4 = 4*pitch, p5 = 5*pitch;
    const int s0  = c[ 2 - p5], s1  = c[   - p5], s2  = c[-2 - p5], s3  = c[-4 - p4];
    const int s4  = c[-5 - p2], s5  = c[-5     ], s6  = c[-5 + p2], s7  = c[-4 + p4];
    const int s8  = c[-2 + p5], s9  = c[     p5], s10 = c[ 2 + p5], s11 = c[ 4 + p4];
    const int s12 = c[ 5 + p2], s13 = c[ 5     ], s14 = c[ 5 - p2], s15 = c[ 4 - p4];
    const int q0 = s0 + s8, q1 = s1 + s9, q2 = s2 + s10, q3 = s3 + s11;
    const int q4 = s4 + s12, q5 = s5 + s13, q6 = s6 + s14, q7 = s7 + s15;
    const int sum  = abs(q0 - q4) + abs(q1 - q5) + abs(q2 - q6) + abs(q3 - q7);
    const int diff = abs(s0 - s8) + abs(s1 - s9) + abs(s2 - s10) + abs(s3 - s11) +
                     abs(s4 - s12) + abs(s5 - s13) + abs(s6 - s14) + abs(s7 - s15);
    const int mean = (q0 + q1 + q2 + q3) + (q4 + q5 + q6 + q7);
    const int local_mean = (c[-1] + c[0] + c[1]) * 16 / 3;
    return sum - diff - abs(mean - local_mean);
}

// L2 arithmetic for four pixels held in byte lanes: s[k] = ring sample k of the four pixels.
// T[p] += sum of chords - sum of diameters of pixel p (|T| <= 2040). The sixteen VABSDIFF4 run on
// the ALU pipe; the per-pixel sums are taken with IDP.4A (dot product with a one-hot +-1 byte
// vector), which runs on the otherwise idle FMA pipe -- the kernel is ALU-pipe bound.
__device__ __forceinline__ int dp4a_us(uint32_t a_u8x4, int b_s8x4, int c)
{
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a_u8x4), "r"(b_s8x4), "r"(c));
    return d;
}
__device__ __forceinline__ void l2_word(const uint32_t (&s)[16], int (&T)[4])
{
#pragma unroll
    for (int i = 0; i < 4; i++)
    {
        const uint32_t u0 = vabs4(s[i], s[i + 4]),  u1 = vabs4(s[i + 8], s[i + 12]);
        const uint32_t v0 = vabs4(s[i], s[i + 8]),  v1 = vabs4(s[i + 4], s[i + 12]);
#pragma unroll
        for (int p = 0; p < 4; p++)
        {
            const int plus = 1 << (8 * p), minus = (int)(0xFFu << (8 * p));
            T[p] = dp4a_us(u0, plus, T[p]);  T[p] = dp4a_us(u1, plus, T[p]);
            T[p] = dp4a_us(v0, minus, T[p]); T[p] = dp4a_us(v1, minus, T[p]);
        }
    }
}

struct StripCtx
{
    const uint8_t* img;     // the frame
    int pitch, w;
    int x0;                 // first pixel of this warp's first 8-pixel cell
    int cell_pitch;         // pixels between the cells of consecutive lanes
    int ys, ye;             // output rows of the segment
    cand_t* out; uint32_t* count; int cap;
};

// L3: exact test of queued pixels, 32 at a time (all of them when `final`).
template<class Q>
__device__ __forceinline__ void l3_phase(const StripCtx& c, Q* q, uint32_t& l3_head, bool final, int lane)
{
    __syncwarp();
    const uint32_t tail = *(volatile uint32_t*)&q->l3_tail;
    while (tail - l3_head >= 32u || (final && tail != l3_head))
    {
        const uint32_t n = min(32u, tail - l3_head);
        int r = 0, x = 0, y = 0;
        if ((uint32_t)lane < n)
        {
            const uint32_t ent = q->l3q[(l3_head + lane) % kL3QCap];
            x = ent & 0xFFFF; y = ent >> 16;
            r = chess_exact(c.img + (size_t)y * c.pitch + x, c.pitch);
        }
        l3_head += n;
        const bool hit = r > kRespMin;
        const uint32_t b = __ballot_sync(kFull, hit);
        if (b)
        {
            uint32_t base = 0;
            if (lane == 0) base = atomicAdd(c.count, (uint32_t)__popc(b));
            base = __shfl_sync(kFull, base, 0);
            if (hit)
            {
                const uint32_t idx = base + __popc(b & ((1u << lane) - 1));
                if (idx < (uint32_t)c.cap) c.out[idx] = cand_pack(x, y, r);
            }
        }
    }
    __syncwarp();
}

// One L2 batch: lane k < n takes unit head+k of the warp's queue, an 8-pixel row cell flagged by
// L1, and reads its 7 ring rows from the staged blocks still resident in shared memory (a cell of
// block B needs rows of blocks B and B-1). unit = one-hot step bit << 5 | lane column; units before
// queue position `mark` belong to block it-1, the others to block it.
// Pixels with T >= 16 go to the L3 queue.
struct L2Ctx
{
    const uint8_t* ring;
    uint32_t off0, off1, off2;   // byte offsets of the stages of blocks it, it-1, it-2 (+ the warp's column base)
    int ybase;                   // first output row of block it
    uint32_t mark;               // queue position of the first unit of block it
    int row_bytes;
};

template<class Q>
__device__ __forceinline__ void l2_batch(const StripCtx& c, const L2Ctx& L, Q* q, uint32_t head, uint32_t n, int lane)
{
    if ((uint32_t)lane < n)
    {
        const uint32_t u = q->l2q[(head + lane) % Q::kL2QCap];
        const int col = u & 31, j = 31 - __clz(u >> 5);
        const bool cur = (int)(head + lane - L.mark) >= 0;
        const uint32_t sA = (cur ? L.off0 : L.off1) + c.cell_pitch * col, sB = (cur ? L.off1 : L.off2) + c.cell_pitch * col;
        const int y = (cur ? L.ybase : L.ybase - kBlkRows) + j, X = c.x0 + c.cell_pitch * col;
        // ring row y+dy was staged (5 - dy) steps before step j: in block B if j >= 5 - dy, else in B-1
        uint2 A[7], B[7], C[7];
        const int behind[7] = { 10, 9, 7, 5, 3, 1, 0 };
#pragma unroll
        for (int k = 0; k < 7; k++)
        {
            const int r = j - behind[k];
            const uint8_t* p = L.ring + (r >= 0 ? sA + r * L.row_bytes : sB + (r + kBlkRows) * L.row_bytes);
            A[k] = *reinterpret_cast<const uint2*>(p);        // X-8, X-4
            B[k] = *reinterpret_cast<const uint2*>(p + 8);    // X,   X+4
            C[k] = *reinterpret_cast<const uint2*>(p + 16);   // X+8, X+12
        }
        uint32_t s[16];
        int T0[4] = { 0, 0, 0, 0 }, T1[4] = { 0, 0, 0, 0 };
        // pixels X..X+3
        s[0]  = __byte_perm(B[0].x, B[0].y, 0x5432); s[1]  = B[0].x; s[2]  = __byte_perm(A[0].y, B[0].x, 0x5432);
        s[3]  = A[1].y;                              s[15] = B[1].y;
        s[4]  = __byte_perm(A[2].x, A[2].y, 0x6543); s[14] = __byte_perm(B[2].y, C[2].x, 0x4321);
        s[5]  = __byte_perm(A[3].x, A[3].y, 0x6543); s[13] = __byte_perm(B[3].y, C[3].x, 0x4321);
        s[6]  = __byte_perm(A[4].x, A[4].y, 0x6543); s[12] = __byte_perm(B[4].y, C[4].x, 0x4321);
        s[7]  = A[5].y;                              s[11] = B[5].y;
        s[8]  = __byte_perm(A[6].y, B[6].x, 0x5432); s[9]  = B[6].x; s[10] = __byte_perm(B[6].x, B[6].y, 0x5432);
        l2_word(s, T0);
        // pixels X+4..X+7: everything one word to the right
        s[0]  = __byte_perm(B[0].y, C[0].x, 0x5432); s[1]  = B[0].y; s[2]  = __byte_perm(B[0].x, B[0].y, 0x5432);
        s[3]  = B[1].x;                              s[15] = C[1].x;
        s[4]  = __byte_perm(A[2].y, B[2].x, 0x6543); s[14] = __byte_perm(C[2].x, C[2].y, 0x4321);
        s[5]  = __byte_perm(A[3].y, B[3].x, 0x6543); s[13] = __byte_perm(C[3].x, C[3].y, 0x4321);
        s[6]  = __byte_perm(A[4].y, B[4].x, 0x6543); s[12] = __byte_perm(C[4].x, C[4].y, 0x4321);
        s[7]  = B[5].x;                              s[11] = C[5].x;
        s[8]  = __byte_perm(B[6].x, B[6].y, 0x5432); s[9]  = B[6].y; s[10] = __byte_perm(B[6].y, C[6].x, 0x5432);
        l2_word(s, T1);

        if (max(max(max(T0[0], T0[1]), max(T0[2], T0[3])), max(max(T1[0], T1[1]), max(T1[2], T1[3]))) >= 16)
        {
            // rare: some of the eight pixels go on to the exact test. bit p of `hits` = pixel X+p
            uint32_t hits = 0;
#pragma unroll
            for (int pp = 0; pp < 4; pp++) hits |= (uint32_t)(T0[pp] >= 16) << pp | (uint32_t)(T1[pp] >= 16) << (4 + pp);
            while (hits)
            {
                const int pp = __ffs(hits) - 1;
                hits &= hits - 1;
                const int x = X + pp;
                if (x >= kMargin && x < c.w - kMargin)
                {
                    const uint32_t pos = atomicAdd(&q->l3_tail, 1u);
                    q->l3q[pos % kL3QCap] = ((uint32_t)y << 16) | (uint32_t)x;
                }
            }
        }
    }
    __syncwarp();
}

template<int NW, bool CARRY>
__global__ void __launch_bounds__(NW * 32, 5)
chess_cascade_kernel(const __grid_constant__ CUtensorMap tmap, FrameSet fs, CascadeParams tp,
                     cand_t* __restrict__ cand, uint32_t* __restrict__ counts)
{
    using G = Geo<NW>;
    extern __shared__ __align__(128) uint8_t smem[];
    // layout: ring[stages][kStageBytes] | full_bar[stages] | released[stages] | queues[NW]
    uint8_t*  ring      = smem;
    uint64_t* full_bar  = reinterpret_cast<uint64_t*>(smem + (size_t)tp.stages * G::kStageBytes);
    uint32_t* released  = reinterpret_cast<uint32_t*>(full_bar + tp.stages);     // warps done with the stage's current block
    using Q = WarpQueues<CARRY>;
    Q* queues = reinterpret_cast<Q*>(released + ((tp.stages + 3) & ~3));

    const int tid = threadIdx.x, lane = tid & 31, wi = tid >> 5;
    const int nst = tp.stages;
    if (tid == 0)
    {
        for (int s = 0; s < nst; s++) { mbar_init(&full_bar[s], 1); released[s] = 0; }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    Q* q = &queues[wi];
    if (lane == 0) { q->l2_tail = 0; q->l3_tail = 0; }
    __syncthreads();

    const int item  = blockIdx.x;
    const int strip = item % tp.nstrips;
    const int seg   = (item / tp.nstrips) % tp.nsegs;
    const int f     = item / (tp.nstrips * tp.nsegs);
    const int xs = strip * G::kStripW;                   // first pixel of the strip
    const int ys = kMargin + seg * tp.seg_rows;          // first output row of the segment
    const int ye = min(ys + tp.seg_rows, fs.h - kMargin);
    // Step j of block `it` stages frame row rbase + it*11 + j: the "+5" row of output row y = that - 5.
    // The first ten staged rows (ys-5 .. ys+4) only prime the register window.
    const int rbase = ys - 5;
    const int nit = (ye - ys + 10 + kBlkRows - 1) / kBlkRows;

    StripCtx c;
    c.img = fs.base + (size_t)f * fs.frame_stride; c.pitch = fs.pitch; c.w = fs.w;
    // The 8-pixel cells of the strip are dealt to the warps round-robin (lane l of warp w owns cell
    // NW*l + w): an edge that crosses the strip then loads every warp alike, which keeps the warps
    // of a CTA, who share the staged rows, in step. 64-bit shared loads at this lane pitch are
    // still bank-conflict free (16 lanes x 24 or 16 bytes cover the 32 banks once).
    c.x0 = xs + wi * kLanePx; c.cell_pitch = NW * kLanePx; c.ys = ys; c.ye = ye;
    c.out = cand + (size_t)f * tp.cap; c.count = counts + f; c.cap = tp.cap;

    // One thread requests block `it` into its stage. The ring is filled once here; after that a
    // stage is refilled by whichever warp is the LAST to be done with it (release() below), so no
    // warp ever waits for a free slot and every copy is in flight as early as it can be.
    auto issue = [&](int it, int s)
    {
        mbar_arrive_expect_tx(&full_bar[s], G::kRowBytes * kBlkRows);
        tma_load_3d(ring + (size_t)s * G::kStageBytes, &tmap, (xs - kCHalo) / 4, rbase + it * kBlkRows, f, &full_bar[s]);
    };
    auto release = [&](int it, int s)    // this warp no longer reads stage s = block `it` (call with the warp converged)
    {
        __syncwarp();
        if (lane == 0)
        {
            __threadfence_block();
            if (atomicAdd(&released[s], 1u) == NW - 1)
            {
                released[s] = 0;
                __threadfence_block();
                if (it + nst < nit)
                {
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    issue(it + nst, s);
                }
            }
        }
    };
    if (tid == 0)
        for (int it = 0; it < nst && it < nit; it++) issue(it, it);

    // byte offset, in a staged row, of this lane's byte X-8 (X = c.x0 + cell_pitch*lane)
    const int lane_off = wi * kLanePx + NW * kLanePx * lane + (kCHalo - 8);

    // Register window (slot = step at which the row was staged). The two words of a lane use mirrored
    // chords so that they share one PRMT'd word per row (bytes X+2..X+5 are dx = +2 for pixels X..X+3
    // and dx = -2 for pixels X+4..X+7):
    //   PP2  = bytes X+2..X+5    rows y-5 and y+5, both words                 live 11 rows
    //   PM5a = bytes X-5..X-2    dx = -5 of rows y-2, y, y+2, pixels X..X+3    live 8 rows
    //   PP5b = bytes X+9..X+12   dx = +5 of rows y-2, y, y+2, pixels X+4..X+7  live 8 rows
    //   the previous row's aligned words Wm1, W0, W1, W2 (ring dx = -4/+4 of row y+4)
    uint32_t PP2[kBlkRows], PM5a[kBlkRows], PP5b[kBlkRows];
#pragma unroll
    for (int j = 0; j < kBlkRows; j++) PP2[j] = PM5a[j] = PP5b[j] = 0;
    uint32_t pWm1 = 0, pW0 = 0, pW1 = 0, pW2 = 0;

    uint32_t q_head = 0, prev_mark = 0, l3_head = 0;
    L2Ctx L;
    L.ring = ring; L.off0 = L.off1 = L.off2 = 0; L.row_bytes = G::kRowBytes;
    // cells without a pixel in [7,w-7) are never tested (their row bytes may lie beyond the pitch)
    const bool lane_valid = c.x0 + c.cell_pitch * lane < fs.w - kMargin;

    int s = 0, s_rel = 0;          // stage of block it / of the block released next
    uint32_t full_parity = 0;
#pragma unroll 1
    for (int it = 0; it < nit; it++)
    {
        mbar_wait_warp(&full_bar[s], full_parity);
        __syncwarp();

        const uint8_t* stage = ring + (size_t)s * G::kStageBytes + lane_off;
        uint32_t flagbits = 0;
#pragma unroll
        for (int j = 0; j < kBlkRows; j++)
        {
            constexpr int N = kBlkRows;
            const uint2 A = *reinterpret_cast<const uint2*>(stage + j * G::kRowBytes);        // X-8, X-4
            const uint2 B = *reinterpret_cast<const uint2*>(stage + j * G::kRowBytes + 8);    // X,   X+4
            const uint2 C = *reinterpret_cast<const uint2*>(stage + j * G::kRowBytes + 16);   // X+8, X+12
            const uint32_t pm5a = __byte_perm(A.x, A.y, 0x6543);
            const uint32_t pp2  = __byte_perm(B.x, B.y, 0x5432);
            const uint32_t pp5b = __byte_perm(C.x, C.y, 0x4321);
            // slots of the rows staged k steps ago
            const int r10 = (j + 1) % N, r7 = (j + 4) % N, r5 = (j + 6) % N, r3 = (j + 8) % N;
            // chords (p in {a,c}, q in {b,d}), output row y = (row staged now) - 5:
            //   pixels X..X+3:    i=0: s0 (+2,-5) | s4 (-5,-2)    i=1: s9 (0,+5) | s5  (-5,0)    i=2: s10 (+2,+5) | s6  (-5,+2)
            //   pixels X+4..X+7:  i=0: s8 (-2,+5) | s12 (+5,+2)   i=1: s9 (0,+5) | s13 (+5,0)    i=2: s2  (-2,-5) | s14 (+5,-2)
            //   both:             i=3: s7 (-4,+4) | s11 (+4,+4)
            const uint32_t u0a = vabs4(PP2[r10], PM5a[r7]),  u0b = vabs4(pp2, PP5b[r3]);
            const uint32_t u1a = vabs4(B.x, PM5a[r5]),       u1b = vabs4(B.y, PP5b[r5]);
            const uint32_t u2a = vabs4(pp2, PM5a[r3]),       u2b = vabs4(PP2[r10], PP5b[r7]);
            const uint32_t u3a = vabs4(pWm1, pW1),           u3b = vabs4(pW0, pW2);
            // packed byte sums: bit 7 of a lane of sa/sb = that pixel's four chords sum to >= 8, PROVIDED no
            // lane can wrap, i.e. every chord <= 31 (4*31 + 0x78 < 256). That is what ha/hb guarantee: the
            // sum of all sixteen chord bytes of a word (IDP.4A against 1,1,1,1 -- FMA pipe, which idles
            // otherwise) is < 32. A word with ha >= 32 holds a chord sum >= 8 anyway and is flagged.
            const uint32_t sa = packed_sum4(u0a, u1a, u2a, u3a);
            const uint32_t sb = packed_sum4(u0b, u1b, u2b, u3b);
            const int ha = dp4a_us(u3a, 0x01010101, dp4a_us(u2a, 0x01010101, dp4a_us(u1a, 0x01010101, dp4a_us(u0a, 0x01010101, 0))));
            const int hb = dp4a_us(u3b, 0x01010101, dp4a_us(u2b, 0x01010101, dp4a_us(u1b, 0x01010101, dp4a_us(u0b, 0x01010101, 0))));
            if ((((sa | sb) & 0x80808080u) | ((uint32_t)(ha | hb) & ~31u)) != 0) flagbits |= 1u << j;
            PP2[j] = pp2; PM5a[j] = pm5a; PP5b[j] = pp5b;
            pWm1 = A.y; pW0 = B.x; pW1 = B.y; pW2 = C.x;
        }

        // rows of this block that belong to the segment: output row of step j is ybase + j
        const int ybase = rbase + it * kBlkRows - 5;
        const int lo = max(ys - ybase, 0), hi = min(ye - ybase, kBlkRows);
        flagbits &= hi > lo && lane_valid ? ((1u << hi) - (1u << lo)) : 0u;

        // one L2 unit per flagged 8-pixel row cell
        if (flagbits)
        {
            // unit = one-hot row bit << 5 | lane column (the consumer finds the row; which block a unit
            // belongs to follows from its position in the queue)
            uint32_t pos = atomicAdd(&q->l2_tail, (uint32_t)__popc(flagbits));
            if (CARRY) pos %= Q::kL2QCap;          // (without carry the queue restarts at 0 every block: no wrap)
            do
            {
                // two cells per round: the loop runs as often as the busiest lane has pairs of flagged rows (+0.5 %; four
                // per round measured the same)
                const uint32_t low = flagbits & (0u - flagbits);
                flagbits ^= low;
                q->l2q[pos] = (uint16_t)(low * 32u + (uint32_t)lane);
                if (CARRY) pos = (pos + 1) % Q::kL2QCap; else ++pos;
                if (flagbits)
                {
                    const uint32_t low2 = flagbits & (0u - flagbits);
                    flagbits ^= low2;
                    q->l2q[pos] = (uint16_t)(low2 * 32u + (uint32_t)lane);
                    if (CARRY) pos = (pos + 1) % Q::kL2QCap; else ++pos;
                }
            } while (flagbits);
        }
        __syncwarp();
        const uint32_t tail = *(volatile uint32_t*)&q->l2_tail;
        L.off2 = L.off1; L.off1 = L.off0; L.off0 = (uint32_t)s * G::kStageBytes + wi * kLanePx + (kCHalo - 8);
        L.ybase = ybase; L.mark = prev_mark;
        // full batches; then whatever is left of the PREVIOUS block (its older stage is about to be handed back)
        while (tail - q_head >= 32u)
        {
            l2_batch(c, L, q, q_head, 32u, lane);
            q_head += 32u;
            l3_phase(c, q, l3_head, false, lane);
        }
        if ((int)(prev_mark - q_head) > 0 || (!CARRY && tail != q_head))
        {
            l2_batch(c, L, q, q_head, tail - q_head, lane);
            q_head = tail;
            l3_phase(c, q, l3_head, false, lane);
        }
        prev_mark = tail;
        if (!CARRY)
        {
            // the queue is empty: the next block's cells start at position 0 again
            __syncwarp();
            if (lane == 0) q->l2_tail = 0;
            q_head = 0; prev_mark = 0;
        }
        // every cell of block it-1 is settled: this warp is done with the stage of block it-2
        // (of block it-1 when cells are never carried over)
        constexpr int hold = CARRY ? 2 : 1;
        if (it >= hold) { release(it - hold, s_rel); if (++s_rel == nst) s_rel = 0; }
        if (++s == nst) { s = 0; full_parity ^= 1; }
    }
    {
        const uint32_t tail = *(volatile uint32_t*)&q->l2_tail;
        while (tail != q_head)
        {
            const uint32_t n = min(32u, tail - q_head);
            l2_batch(c, L, q, q_head, n, lane);
            q_head += n;
            l3_phase(c, q, l3_head, false, lane);
        }
    }
    l3_phase(c, q, l3_head, true, lane);
}

bool make_cascade_map(CUtensorMap* map, const FrameSet& fs, int row_bytes)
{
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) return false;
    if (((uintptr_t)fs.base & 15) || (fs.pitch & 15)) return false;
    size_t fstride = fs.frame_stride;
    if (fs.nframes == 1) fstride = ((size_t)fs.pitch * fs.h + 15) & ~(size_t)15;
    if (fstride & 15) return false;
    // the row is declared ceil(w/4) words wide, not `pitch`: whatever lies beyond is zero-filled by TMA instead of
    // fetched (a frame may be a crop of a larger allocation whose last row ends before base + h*pitch). The up to
    // three bytes between w and the end of the last word are inside the row's pitch (pitch % 16 == 0 >= w) and only
    // ever feed pixels outside [7,w-7).
    const cuuint64_t dims[3]    = { (cuuint64_t)((fs.w + 3) / 4), (cuuint64_t)fs.h, (cuuint64_t)fs.nframes };
    const cuuint64_t strides[2] = { (cuuint64_t)fs.pitch, (cuuint64_t)fstride };
    const cuuint32_t box[3]     = { (cuuint32_t)(row_bytes / 4), kBlkRows, 1 };
    const cuuint32_t estr[3]    = { 1, 1, 1 };
    const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, (void*)fs.base, dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

template<int NW, bool CARRY>
cudaError_t launch_nw(const FrameSet& fs, CascadeParams tp, cand_t* cand, uint32_t* counts, cudaStream_t stream, bool* ok)
{
    using G = Geo<NW>;
    CUtensorMap map;
    *ok = make_cascade_map(&map, fs, G::kRowBytes);
    if (!*ok) return cudaSuccess;
    const size_t smem = (size_t)tp.stages * G::kStageBytes + tp.stages * sizeof(uint64_t) + ((tp.stages + 3) & ~3) * sizeof(uint32_t) +
                        NW * sizeof(WarpQueues<CARRY>);
    if (smem > 48 * 1024)
    {
        // per device, idempotent and cheap: set on every launch rather than tracking devices
        cudaError_t e = cudaFuncSetAttribute(chess_cascade_kernel<NW, CARRY>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    const long long items = (long long)fs.nframes * tp.nstrips * tp.nsegs;
    if (items > 0x7fffffffLL) return cudaErrorInvalidValue;
    chess_cascade_kernel<NW, CARRY><<<(unsigned)items, NW * 32, smem, stream>>>(map, fs, tp, cand, counts);
    return cudaGetLastError();
}

int env_int(const char* name, int dflt, int lo, int hi)
{
    if (const char* e = getenv(name)) { const int v = atoi(e); if (v >= lo && v <= hi) return v; }
    return dflt;
}
}   // namespace

// Returns cudaSuccess with *launched = false when the frames do not meet TMA's alignment rules
// (the caller then uses the tiled kernel's cooperative loader).
cudaError_t launch_chess_sparse_cascade(const FrameSet& fs, cand_t* cand, uint32_t* counts,
                                        int cand_capacity, cudaStream_t stream, bool* launched)
{
    *launched = true;
    if (fs.w <= 2*kMargin || fs.h <= 2*kMargin || fs.nframes <= 0) return cudaSuccess;
    if (fs.h > 32767 || fs.w > 32767) { *launched = false; return cudaSuccess; }

    // strip width: three warps per CTA (768-pixel strips, 4 % halo) unless that stages over 30 % more bytes
    // than one-warp strips would (narrow frames). Two-warp strips measured slower than either at every
    // size tried (fewer warps per SM for the shared memory they hold); MRG_B200_K1_NW forces a width.
    int nw = env_int("MRG_B200_K1_NW", 0, 1, 3);
    if (nw == 0)
    {
        auto staged = [&](int k) { const int sw = kWarpPx * k; return (long long)((fs.w - kMargin + sw - 1) / sw) * (sw + 2 * kCHalo); };
        nw = staged(3) * 10 <= staged(1) * 13 ? 3 : 1;
    }
    CascadeParams tp;
    tp.cap = cand_capacity;
    // a stage is handed back two blocks after it was consumed (L2 reads it), so 3 stages are always held
    // tuning knobs (defaults are what bench.py measures): MRG_B200_K1_NW, MRG_B200_K1_STAGES, MRG_B200_K1_NOCARRY
    const int nocarry = env_int("MRG_B200_K1_NOCARRY", 0, 0, 1);
    tp.stages = env_int("MRG_B200_K1_STAGES", nocarry ? 3 : 4, nocarry ? 3 : 4, 12);
    const int sw = kWarpPx * nw;
    tp.nstrips = (fs.w - kMargin + sw - 1) / sw;
    const int out_rows = fs.h - 2*kMargin;
    // Work items = (frame, strip, row segment). Many short segments balance better (a CTA's time depends on
    // how many edges cross its segment, and the last wave of CTAs is shorter) but each segment re-reads 10
    // rows to prime its register window: ~275-row segments measured best on 4K batches (0.31 -> 0.35 of HBM
    // peak against one segment per strip). Small batches get more, shorter segments to fill the chip
    // (>= 2368 items), never shorter than 44 rows.
    const int target_rows = env_int("MRG_B200_K1_SEGROWS", 275, 44, 1 << 20);
    const long long want_items = env_int("MRG_B200_K1_ITEMS", 148 * 4 * 4, 1, 1 << 24);
    const long long per_seg = (long long)fs.nframes * tp.nstrips;
    long long nsegs = std::max((want_items + per_seg - 1) / per_seg, (long long)((out_rows + target_rows - 1) / target_rows));
    int seg_rows = (int)((out_rows + nsegs - 1) / nsegs);
    seg_rows = ((seg_rows + kBlkRows - 1) / kBlkRows) * kBlkRows;
    if (seg_rows < 44) seg_rows = 44;
    tp.seg_rows = seg_rows;
    tp.nsegs = (out_rows + seg_rows - 1) / seg_rows;

    if (nw == 1) return nocarry ? launch_nw<1, false>(fs, tp, cand, counts, stream, launched) : launch_nw<1, true>(fs, tp, cand, counts, stream, launched);
    if (nw == 2) return nocarry ? launch_nw<2, false>(fs, tp, cand, counts, stream, launched) : launch_nw<2, true>(fs, tp, cand, counts, stream, launched);
    return nocarry ? launch_nw<3, false>(fs, tp, cand, counts, stream, launched) : launch_nw<3, true>(fs, tp, cand, counts, stream, launched);
}

}
