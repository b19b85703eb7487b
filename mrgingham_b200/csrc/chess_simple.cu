// The straightforward kernels: pyramid level (K0), dense ChESS response (the ChESS_response_5
// API) and the simple one-thread-per-pixel sparse ChESS kernel that the tiled kernel
// (chess_tiled.cu) is cross-checked against. Integer arithmetic only.
#include "kernels.cuh"

namespace mrgb200
{

// ------------------------------------------------------------------------------------------------
// ChESS response of one pixel, ring radius 5. Semantics of ChESS.c:62-105:
//   16 ring samples s0..s15, opposite pairs (s_k, s_k+8);
//   sum  = sum_i |(s_i + s_i+8) - (s_i+4 + s_i+12)|     i = 0..3
//   diff = sum_k |s_k - s_k+8|                           k = 0..7
//   mean = sum of the 16 samples;  local_mean = (I[x-1]+I[x]+I[x+1])*16/3 (truncating)
//   response = sum - diff - |mean - local_mean|
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int chess_response_at(const uint8_t* __restrict__ c, int pitch)
{
    // c points at the pixel; ring offsets (dx,dy) as in ChESS.c:68-83
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

// ------------------------------------------------------------------------------------------------
// K1 simple: thread per pixel of the interior; warp-aggregated append of {r > 15}
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
chess_sparse_simple_kernel(FrameSet fs, cand_t* __restrict__ cand, uint32_t* __restrict__ counts, int cap)
{
    const int f = blockIdx.z;
    const int x = kMargin + blockIdx.x * 32 + (threadIdx.x & 31);
    const int y = kMargin + blockIdx.y * 8  + (threadIdx.x >> 5);
    const bool inside = x < fs.w - kMargin && y < fs.h - kMargin;

    int r = 0;
    if (inside)
        r = chess_response_at(fs.base + (size_t)f * fs.frame_stride + (size_t)y * fs.pitch + x, fs.pitch);

    const bool hit = r > kRespMin;
    const unsigned ballot = __ballot_sync(0xffffffffu, hit);
    if (ballot == 0) return;
    const int lane = threadIdx.x & 31;
    unsigned base = 0;
    if (lane == 0) base = atomicAdd(&counts[f], (unsigned)__popc(ballot));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (hit)
    {
        const unsigned idx = base + __popc(ballot & ((1u << lane) - 1));
        if (idx < (unsigned)cap) cand[(size_t)f * cap + idx] = cand_pack(x, y, r);
    }
}

cudaError_t launch_chess_sparse_simple(const FrameSet& fs, cand_t* cand, uint32_t* counts,
                                       int cand_capacity, cudaStream_t stream)
{
    if (fs.w <= 2*kMargin || fs.h <= 2*kMargin || fs.nframes <= 0) return cudaSuccess;
    dim3 grid((fs.w - 2*kMargin + 31) / 32, (fs.h - 2*kMargin + 7) / 8, fs.nframes);
    chess_sparse_simple_kernel<<<grid, 256, 0, stream>>>(fs, cand, counts, cand_capacity);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// dense response: interior only, border elements of `response` are never written
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
chess_dense_kernel(FrameSet fs, int16_t* __restrict__ response, size_t resp_frame_stride)
{
    const int f = blockIdx.z;
    const int x = kMargin + blockIdx.x * 32 + (threadIdx.x & 31);
    const int y = kMargin + blockIdx.y * 8  + (threadIdx.x >> 5);
    if (x >= fs.w - kMargin || y >= fs.h - kMargin) return;
    const int r = chess_response_at(fs.base + (size_t)f * fs.frame_stride + (size_t)y * fs.pitch + x, fs.pitch);
    response[(size_t)f * resp_frame_stride + (size_t)y * fs.w + x] = (int16_t)r;
}

cudaError_t launch_chess_dense(const FrameSet& fs, int16_t* response, size_t response_frame_stride_elems,
                               cudaStream_t stream)
{
    if (fs.w <= 2*kMargin || fs.h <= 2*kMargin || fs.nframes <= 0) return cudaSuccess;
    dim3 grid((fs.w - 2*kMargin + 31) / 32, (fs.h - 2*kMargin + 7) / 8, fs.nframes);
    chess_dense_kernel<<<grid, 256, 0, stream>>>(fs, response, response_frame_stride_elems);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// K0 pyramid level. Exact integer model of cv::resize(src, dst, Size(), 1/2^L, 1/2^L,
// INTER_LINEAR) (find_chessboard_corners.cc:449-450), pinned against cv2 in
// tests/test_pyramid_model.py:
//   B = 2^L; out[dy][dx] = (I[y0][x0] + I[y0][x1] + I[y1][x0] + I[y1][x1] + 2) >> 2,
//   x0 = min(B*dx + B/2 - 1, W-1), x1 = min(x0+1, W-1), same in y;
//   for L == 1 a trailing partial cell (W or H == 3 mod 4) is the round-half-even mean of the
//   pixels that exist.
// Every level is computed from level 0 directly, never cascaded.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int rint_half(int s) { const int q = s >> 1; return q + ((s & 1) & (q & 1)); }

__global__ void __launch_bounds__(256)
pyramid_kernel(FrameSet src, int level, uint8_t* __restrict__ dst, int dst_pitch, size_t dst_frame_stride,
               int ow, int oh)
{
    const int f  = blockIdx.z;
    const int dx = blockIdx.x * 32 + (threadIdx.x & 31);
    const int dy = blockIdx.y * 8  + (threadIdx.x >> 5);
    if (dx >= ow || dy >= oh) return;
    const uint8_t* img = src.base + (size_t)f * src.frame_stride;
    const int B = 1 << level, W = src.w, H = src.h;
    const int x0 = min(B*dx + B/2 - 1, W-1), x1 = min(x0 + 1, W-1);
    const int y0 = min(B*dy + B/2 - 1, H-1), y1 = min(y0 + 1, H-1);
    int v = (img[(size_t)y0*src.pitch + x0] + img[(size_t)y0*src.pitch + x1] +
             img[(size_t)y1*src.pitch + x0] + img[(size_t)y1*src.pitch + x1] + 2) >> 2;
    if (level == 1)
    {
        const bool px = 2*ow > W && dx == ow-1;
        const bool py = 2*oh > H && dy == oh-1;
        if (px && py)
            v = img[(size_t)(H-1)*src.pitch + W-1];
        else if (px)
        {
            const int ya = 2*dy, yb = min(ya+1, H-1);
            v = rint_half(img[(size_t)ya*src.pitch + W-1] + img[(size_t)yb*src.pitch + W-1]);
        }
        else if (py)
        {
            const int xa = 2*dx, xb = min(xa+1, W-1);
            v = rint_half(img[(size_t)(H-1)*src.pitch + xa] + img[(size_t)(H-1)*src.pitch + xb]);
        }
    }
    dst[(size_t)f * dst_frame_stride + (size_t)dy * dst_pitch + dx] = (uint8_t)v;
}

// ------------------------------------------------------------------------------------------------
// Kpre: the box blur the reference CLI applies before the detector by default
// (mrgingham-from-image.cc:106-111, --blur 1): cv::blur(Size(1+2R,1+2R)), BORDER_REFLECT_101,
// = floor((sum + (k*k-1)/2) / (k*k)) for 8-bit data. A CTA blurs a 128 x 8 output tile from a
// shared-memory tile with the halo; 2 bytes of HBM traffic per pixel (1 read + 1 written).
// ------------------------------------------------------------------------------------------------
constexpr int kBlurMaxR = 4, kBlurTW = 128;
__device__ __forceinline__ int reflect101(int i, int n)
{
    if (n == 1) return 0;
    while (i < 0 || i >= n) i = i < 0 ? -i : 2 * (n - 1) - i;
    return i;
}
// Any radius 1..4, any alignment: a CTA owns a 128 x 32 tile. The tile and its halo go to shared memory (rows read
// as words where the frame allows it, reflected at the image border), then the box sum is taken separably: horizontal
// running sums of every staged row into 16-bit lanes, vertical sums of those. 2k additions per pixel instead of k*k.
constexpr int kBlurTH2 = 32;
__global__ void __launch_bounds__(256)
box_blur_kernel(FrameSet fs, int radius, uint8_t* __restrict__ dst, int dst_pitch, size_t dst_frame_stride)
{
    __shared__ uint8_t  tile[kBlurTH2 + 2 * kBlurMaxR][kBlurTW + 2 * kBlurMaxR + 8];
    __shared__ uint16_t hs[kBlurTH2 + 2 * kBlurMaxR][kBlurTW];
    const int f = blockIdx.z, x0 = blockIdx.x * kBlurTW, y0 = blockIdx.y * kBlurTH2;
    const uint8_t* img = fs.base + (size_t)f * fs.frame_stride;
    const int tw = kBlurTW + 2 * radius, th = kBlurTH2 + 2 * radius;
    // stage: one warp per row at a time, lanes along the row (coalesced); reflection only where the tile leaves the image
    for (int ty = threadIdx.x >> 5; ty < th; ty += 8)
    {
        const uint8_t* row = img + (size_t)reflect101(y0 + ty - radius, fs.h) * fs.pitch;
        for (int tx = threadIdx.x & 31; tx < tw; tx += 32)
        {
            const int x = x0 + tx - radius;
            tile[ty][tx] = row[(x >= 0 && x < fs.w) ? x : reflect101(x, fs.w)];
        }
    }
    __syncthreads();
    const int k = 2 * radius + 1, k2 = k * k, half = (k2 - 1) / 2;
    // horizontal sums: thread = 4 adjacent pixels of a staged row
    for (int i = threadIdx.x; i < th * (kBlurTW / 4); i += 256)
    {
        const int ty = i / (kBlurTW / 4), tx = (i % (kBlurTW / 4)) * 4;
        int s = 0;
        for (int dx = 0; dx < k; dx++) s += tile[ty][tx + dx];
        hs[ty][tx] = (uint16_t)s;
#pragma unroll
        for (int p2 = 1; p2 < 4; p2++) { s += tile[ty][tx + p2 - 1 + k] - tile[ty][tx + p2 - 1]; hs[ty][tx + p2] = (uint16_t)s; }
    }
    __syncthreads();
    // vertical sums: thread = one column of 16 rows, sliding
    const int cx = threadIdx.x & 127, half_rows = (threadIdx.x >> 7) * (kBlurTH2 / 2);
    const int x = x0 + cx;
    if (x >= fs.w) return;
    int s = 0;
    for (int dy = 0; dy < k; dy++) s += hs[half_rows + dy][cx];
    uint8_t* o = dst + (size_t)f * dst_frame_stride + x;
    for (int r = 0; r < kBlurTH2 / 2; r++)
    {
        const int y = y0 + half_rows + r;
        if (y >= fs.h) break;
        o[(size_t)y * dst_pitch] = (uint8_t)((s + half) / k2);
        if (r + 1 < kBlurTH2 / 2) s += hs[half_rows + r + k][cx] - hs[half_rows + r][cx];
    }
}

// Fast path for the CLI's default R = 1 (3x3): HBM-bound, 1 byte read + 1 byte written per pixel.
// A lane owns 16 adjacent pixels (one 128-bit load and one 128-bit store per row) and walks down a
// row segment keeping the unpacked 16-bit lanes of the previous two rows in registers: vertical
// 3-sums first (packed IADD3), then horizontal 3-sums with PRMT-shifted copies, neighbours across
// lanes by shuffle. Lanes 0 and 31 of a warp are halo lanes (they load and sum but do not store), so
// a warp emits 480 pixels per row and no lane ever needs data from another warp.
// Needs 16-byte aligned rows on both sides and w % 16 == 0 (every BASELINE size); otherwise the
// generic kernel above runs. floor((s + 4) / 9) = ((s + 4) * 7282) >> 16 for s <= 2295 (checked exhaustively).
constexpr int kBlur3SegRows = 64;
__global__ void __launch_bounds__(128)
box_blur3_kernel(FrameSet fs, uint8_t* __restrict__ dst, int dst_pitch, size_t dst_frame_stride, int strips_per_row)
{
    const int lane = threadIdx.x & 31, wglobal = blockIdx.x * 4 + (threadIdx.x >> 5);
    const int strip = wglobal % strips_per_row, seg = wglobal / strips_per_row, f = blockIdx.y;
    const int w = fs.w, h = fs.h;
    const int x0 = strip * 480 + 16 * (lane - 1);
    const int ys = seg * kBlur3SegRows, ye = min(ys + kBlur3SegRows, h);
    if (ys >= h) return;
    const bool inside = x0 >= 0 && x0 < w;
    const uint8_t* img = fs.base + (size_t)f * fs.frame_stride + (inside ? x0 : 0);
    uint8_t* out = dst + (size_t)f * dst_frame_stride + (inside ? x0 : 0);
    const bool store = inside && lane >= 1 && lane <= 30;

    auto load_row = [&](int y, uint32_t (&u)[8])
    {
        // row y (reflected at the top/bottom edge) as eight words of two 16-bit lanes
        const int yy = h == 1 ? 0 : (y < 0 ? -y : (y >= h ? 2 * (h - 1) - y : y));
        uint4 q = make_uint4(0, 0, 0, 0);
        if (inside) q = *reinterpret_cast<const uint4*>(img + (size_t)yy * fs.pitch);
        u[0] = __byte_perm(q.x, 0, 0x4140); u[1] = __byte_perm(q.x, 0, 0x4342);
        u[2] = __byte_perm(q.y, 0, 0x4140); u[3] = __byte_perm(q.y, 0, 0x4342);
        u[4] = __byte_perm(q.z, 0, 0x4140); u[5] = __byte_perm(q.z, 0, 0x4342);
        u[6] = __byte_perm(q.w, 0, 0x4140); u[7] = __byte_perm(q.w, 0, 0x4342);
    };
    uint32_t ra[8], rb[8], rc[8];
    load_row(ys - 1, ra);
    load_row(ys, rb);
#pragma unroll 1
    for (int y = ys; y < ye; y++)
    {
        load_row(y + 1, rc);
        uint32_t v[8];
#pragma unroll
        for (int k = 0; k < 8; k++) v[k] = ra[k] + rb[k] + rc[k];           // vertical sums, <= 765 per lane
        // neighbours: v[-1] from the lane on the left, v[16] from the lane on the right; at the image's
        // left/right edge BORDER_REFLECT_101 takes v[1] / v[14]
        uint32_t lv = __shfl_up_sync(0xffffffffu, v[7], 1) >> 16;
        uint32_t rv = __shfl_down_sync(0xffffffffu, v[0], 1) & 0xFFFFu;
        if (x0 == 0) lv = v[0] >> 16;
        if (x0 + 16 == w) rv = v[7] & 0xFFFFu;
        uint32_t o[4];
#pragma unroll
        for (int k = 0; k < 8; k++)
        {
            const uint32_t left  = k == 0 ? (lv | (v[0] << 16)) : __byte_perm(v[k - 1], v[k], 0x5432);     // [v(2k-1), v(2k)]
            const uint32_t right = k == 7 ? ((v[7] >> 16) | (rv << 16)) : __byte_perm(v[k], v[k + 1], 0x5432); // [v(2k+1), v(2k+2)]
            const uint32_t s = v[k] + left + right;                                                        // <= 2295 per lane
            const uint32_t q0 = (((s & 0xFFFFu) + 4u) * 7282u) >> 16, q1 = (((s >> 16) + 4u) * 7282u) >> 16;
            const uint32_t pair = q0 | (q1 << 8);
            if (k & 1) o[k >> 1] |= pair << 16; else o[k >> 1] = pair;
        }
        if (store) *reinterpret_cast<uint4*>(out + (size_t)y * dst_pitch) = make_uint4(o[0], o[1], o[2], o[3]);
#pragma unroll
        for (int k = 0; k < 8; k++) { ra[k] = rb[k]; rb[k] = rc[k]; }
    }
}

cudaError_t launch_box_blur(const FrameSet& src, int radius, uint8_t* dst, int dst_pitch, size_t dst_frame_stride, cudaStream_t stream)
{
    if (src.w <= 0 || src.h <= 0 || src.nframes <= 0) return cudaSuccess;
    if (radius < 1 || radius > kBlurMaxR) return cudaErrorInvalidValue;
    const bool aligned = !((uintptr_t)src.base & 15) && !(src.pitch & 15) && !(src.frame_stride & 15) &&
                         !((uintptr_t)dst & 15) && !(dst_pitch & 15) && !(dst_frame_stride & 15) && !(src.w & 15);
    if (radius == 1 && aligned && src.nframes <= 65535)
    {
        const int strips = (src.w + 479) / 480, segs = (src.h + kBlur3SegRows - 1) / kBlur3SegRows;
        dim3 grid((strips * segs + 3) / 4, src.nframes);
        box_blur3_kernel<<<grid, 128, 0, stream>>>(src, dst, dst_pitch, dst_frame_stride, strips);
        return cudaGetLastError();
    }
    dim3 grid((src.w + kBlurTW - 1) / kBlurTW, (src.h + kBlurTH2 - 1) / kBlurTH2, src.nframes);
    box_blur_kernel<<<grid, 256, 0, stream>>>(src, radius, dst, dst_pitch, dst_frame_stride);
    return cudaGetLastError();
}

// Level 1 fast path (every BASELINE size: W % 32 == 0, H even, 16-byte aligned rows): HBM-bound,
// 1 byte read + 1/4 byte written per frame pixel. A thread turns two 128-bit loads (16 pixels of rows
// 2dy and 2dy+1) into 8 output pixels; each 2x2 sum is two IDP.4A against a two-byte mask, seeded with
// the rounding constant: out = (a + b + c + d + 2) >> 2, which is what cv::resize gives for exact halving
// (no partial cells exist when W % 4 != 3 and H % 4 != 3).
__global__ void __launch_bounds__(256)
pyramid_l1_kernel(FrameSet src, uint8_t* __restrict__ dst, int dst_pitch, size_t dst_frame_stride, int ow, int oh)
{
    const int f = blockIdx.z, dy = blockIdx.y * 8 + (threadIdx.x >> 5);
    const int g = blockIdx.x * 32 + (threadIdx.x & 31);          // group of 8 output pixels
    if (dy >= oh || g * 8 >= ow) return;
    const uint8_t* r0 = src.base + (size_t)f * src.frame_stride + (size_t)(2 * dy) * src.pitch + g * 16;
    const uint4 a = *reinterpret_cast<const uint4*>(r0), b = *reinterpret_cast<const uint4*>(r0 + src.pitch);
    const uint32_t wa[4] = { a.x, a.y, a.z, a.w }, wb[4] = { b.x, b.y, b.z, b.w };
    uint32_t o[2] = { 0, 0 };
#pragma unroll
    for (int k = 0; k < 4; k++)
    {
        const uint32_t lo = (uint32_t)__dp4a(wa[k], 0x00000101u, __dp4a(wb[k], 0x00000101u, 2u)) >> 2;
        const uint32_t hi = (uint32_t)__dp4a(wa[k], 0x01010000u, __dp4a(wb[k], 0x01010000u, 2u)) >> 2;
        o[k >> 1] |= (lo | (hi << 8)) << (16 * (k & 1));
    }
    *reinterpret_cast<uint2*>(dst + (size_t)f * dst_frame_stride + (size_t)dy * dst_pitch + g * 8) = make_uint2(o[0], o[1]);
}

// Level 2 fast path (W % 64 == 0, H % 4 == 0): output (dx, dy) averages bytes 1, 2 of the aligned word
// 4dx.. in rows 4dy+1 and 4dy+2; a thread makes 4 outputs from two 128-bit loads. Half of the frame's
// rows are never touched.
__global__ void __launch_bounds__(256)
pyramid_l2_kernel(FrameSet src, uint8_t* __restrict__ dst, int dst_pitch, size_t dst_frame_stride, int ow, int oh)
{
    const int f = blockIdx.z, dy = blockIdx.y * 8 + (threadIdx.x >> 5);
    const int g = blockIdx.x * 32 + (threadIdx.x & 31);          // group of 4 output pixels
    if (dy >= oh || g * 4 >= ow) return;
    const uint8_t* r0 = src.base + (size_t)f * src.frame_stride + (size_t)(4 * dy + 1) * src.pitch + g * 16;
    const uint4 a = *reinterpret_cast<const uint4*>(r0), b = *reinterpret_cast<const uint4*>(r0 + src.pitch);
    const uint32_t wa[4] = { a.x, a.y, a.z, a.w }, wb[4] = { b.x, b.y, b.z, b.w };
    uint32_t o = 0;
#pragma unroll
    for (int k = 0; k < 4; k++)
        o |= ((uint32_t)__dp4a(wa[k], 0x00010100u, __dp4a(wb[k], 0x00010100u, 2u)) >> 2) << (8 * k);
    *reinterpret_cast<uint32_t*>(dst + (size_t)f * dst_frame_stride + (size_t)dy * dst_pitch + g * 4) = o;
}

cudaError_t launch_pyramid(const FrameSet& src, int level, uint8_t* dst, int dst_pitch,
                           size_t dst_frame_stride, int ow, int oh, cudaStream_t stream)
{
    if (ow <= 0 || oh <= 0 || src.nframes <= 0) return cudaSuccess;
    if (level == 2 && !(src.w & 63) && !(src.h & 3) && !((uintptr_t)src.base & 15) && !(src.pitch & 15) && !(src.frame_stride & 15) &&
        !((uintptr_t)dst & 3) && !(dst_pitch & 3) && !(dst_frame_stride & 3) && src.nframes <= 65535)
    {
        dim3 grid((ow / 4 + 31) / 32, (oh + 7) / 8, src.nframes);
        pyramid_l2_kernel<<<grid, 256, 0, stream>>>(src, dst, dst_pitch, dst_frame_stride, ow, oh);
        return cudaGetLastError();
    }
    if (level == 1 && !(src.w & 31) && !(src.h & 1) && !((uintptr_t)src.base & 15) && !(src.pitch & 15) && !(src.frame_stride & 15) &&
        !((uintptr_t)dst & 7) && !(dst_pitch & 7) && !(dst_frame_stride & 7) && src.nframes <= 65535)
    {
        dim3 grid((ow / 8 + 31) / 32, (oh + 7) / 8, src.nframes);
        pyramid_l1_kernel<<<grid, 256, 0, stream>>>(src, dst, dst_pitch, dst_frame_stride, ow, oh);
        return cudaGetLastError();
    }
    dim3 grid((ow + 31) / 32, (oh + 7) / 8, src.nframes);
    pyramid_kernel<<<grid, 256, 0, stream>>>(src, level, dst, dst_pitch, dst_frame_stride, ow, oh);
    return cudaGetLastError();
}


// ---- gather of scattered, equally-sized device images into one batch (mrg_b200_find_corners_mixed_batch) ----
__global__ void __launch_bounds__(256)
gather_frames_kernel(const GatherSrc* __restrict__ srcs, int rows, int cols, uint8_t* __restrict__ dst, size_t dst_pitch, size_t dst_frame_stride)
{
    const int f = blockIdx.z, y = blockIdx.y;
    const GatherSrc s = srcs[f];
    const uint8_t* in = s.data + (size_t)y * s.pitch;
    uint8_t* out = dst + (size_t)f * dst_frame_stride + (size_t)y * dst_pitch;
    const int x16 = (blockIdx.x * blockDim.x + threadIdx.x) * 16;
    if (x16 >= cols) return;
    if ((((uintptr_t)in | s.pitch) & 15) == 0 && x16 + 16 <= cols)
        *reinterpret_cast<uint4*>(out + x16) = *reinterpret_cast<const uint4*>(in + x16);      // dst rows are 16-byte aligned
    else
        for (int i = 0; i < 16 && x16 + i < cols; i++) out[x16 + i] = in[x16 + i];
}

cudaError_t launch_gather_frames(const GatherSrc* srcs, int n, int rows, int cols, uint8_t* dst, size_t dst_pitch,
                                 size_t dst_frame_stride, cudaStream_t stream)
{
    if (n <= 0 || rows <= 0 || cols <= 0) return cudaSuccess;
    const int per_row = (cols + 15) / 16;
    for (int f0 = 0; f0 < n; f0 += 65535)
    {
        const int m = std::min(65535, n - f0);
        gather_frames_kernel<<<dim3((per_row + 255) / 256, rows, m), 256, 0, stream>>>(srcs + f0, rows, cols, dst + (size_t)f0 * dst_frame_stride,
                                                                                        dst_pitch, dst_frame_stride);
    }
    return cudaGetLastError();
}
}
