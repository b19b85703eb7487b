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

cudaError_t launch_pyramid(const FrameSet& src, int level, uint8_t* dst, int dst_pitch,
                           size_t dst_frame_stride, int ow, int oh, cudaStream_t stream)
{
    if (ow <= 0 || oh <= 0 || src.nframes <= 0) return cudaSuccess;
    dim3 grid((ow + 31) / 32, (oh + 7) / 8, src.nframes);
    pyramid_kernel<<<grid, 256, 0, stream>>>(src, level, dst, dst_pitch, dst_frame_stride, ow, oh);
    return cudaGetLastError();
}

}
