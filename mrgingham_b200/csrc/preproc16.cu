// The reference CLI's handling of 16-bit images (mrgingham-from-image.cc:83-93):
//     if --clahe:  cv::normalize(image0, image0, 0, 65535, NORM_MINMAX);  clahe->apply(image0, image0);   // clipLimit 8
//     image0.convertTo(image1, CV_8U, 255./65535.);
// after which the 8-bit path (blur, detector) takes over. OpenCV's arithmetic (third party; pinned to cv2 4.13.0 by
// tests/test_preproc16.py through the CPU restatement) is reproduced exactly:
//   convertTo 16u -> 8u     rint((float)v * (float)(255/65535)), saturated
//   normalize 16u -> 16u    scale = 65535 / (max - min), shift = -min * scale in double; rint(fmaf(v, (float)scale, (float)shift))
//   CLAHE on 16 bits        the 8-bit algorithm with 65536 bins per tile: integer clip + redistribute rule, running
//                           sums scaled by (float)(65535 / tileArea), four table look-ups per pixel and OpenCV's float
//                           bilinear expression with every operation rounded separately.
// Kernels (the 65536-bin histograms and tables live in global memory: 16 MB + 8 MB per frame):
//   Q1 minmax16_kernel      per-frame min / max (atomics)
//   Q2 hist16_kernel        one CTA per (frame, tile, row slab): global atomics on the tile's histogram of the
//                           NORMALISED values (the normalisation is recomputed per pixel, never stored)
//   Q3 lut16_kernel         one CTA per (frame, tile): clip, redistribute, block scan over 65536 bins -> uint16 table
//   Q4 apply16_kernel       per pixel: normalise, four look-ups, blend, then the 8-bit conversion; 1 byte written
//   Q0 convert16to8_kernel  the chain without --clahe
#include <cuda_runtime.h>
#include <float.h>
#include <algorithm>

#include "kernels.cuh"

namespace mrgb200
{
namespace
{
constexpr int kTiles = 8, kHist16 = 65536;

struct Geom16
{
    const uint16_t* base; size_t frame_stride_elems; int pitch_elems;      // source frames
    int w, h, n;
    int tw, th, clip;
    float lut_scale, inv_tw, inv_th;
};

__device__ __forceinline__ int reflect101(int i, int n)
{
    if (n == 1) return 0;
    while (i < 0 || i >= n) { if (i < 0) i = -i; else i = 2 * (n - 1) - i; }
    return i;
}

__device__ __forceinline__ uint8_t to8(float v16)
{
    // convertTo(CV_8U, 255./65535.): float multiply (one rounding), round half to even, saturate
    const int q = __float2int_rn(__fmul_rn(v16, (float)(255. / 65535.)));
    return (uint8_t)min(max(q, 0), 255);
}

__global__ void __launch_bounds__(256)
convert16to8_kernel(Geom16 g, uint8_t* __restrict__ dst, int dst_pitch, size_t dst_frame_stride)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, f = blockIdx.z;
    if (x >= g.w) return;
    const uint16_t v = g.base[(size_t)f * g.frame_stride_elems + (size_t)y * g.pitch_elems + x];
    dst[(size_t)f * dst_frame_stride + (size_t)y * dst_pitch + x] = to8((float)v);
}

__global__ void __launch_bounds__(256)
minmax16_kernel(Geom16 g, unsigned* __restrict__ mm)
{
    const int f = blockIdx.y;
    unsigned mn = 65535u, mx = 0u;
    for (int y = blockIdx.x; y < g.h; y += gridDim.x)
    {
        const uint16_t* row = g.base + (size_t)f * g.frame_stride_elems + (size_t)y * g.pitch_elems;
        for (int x = threadIdx.x; x < g.w; x += blockDim.x) { const unsigned v = row[x]; mn = min(mn, v); mx = max(mx, v); }
    }
    mn = __reduce_min_sync(0xffffffffu, mn); mx = __reduce_max_sync(0xffffffffu, mx);
    if ((threadIdx.x & 31) == 0) { atomicMin(&mm[2 * f], mn); atomicMax(&mm[2 * f + 1], mx); }
}

// scale / shift of cv::normalize(0, 65535, NORM_MINMAX) as the floats convertTo applies
__device__ __forceinline__ void norm_coeffs(const unsigned* __restrict__ mm, int f, float* a, float* b)
{
    const double smin = (double)mm[2 * f], smax = (double)mm[2 * f + 1];
    const double scale = __dmul_rn(65535.0, (__dsub_rn(smax, smin) > DBL_EPSILON ? __ddiv_rn(1.0, __dsub_rn(smax, smin)) : 0.0));
    const double shift = __dsub_rn(0.0, __dmul_rn(smin, scale));
    *a = (float)scale; *b = (float)shift;
}
__device__ __forceinline__ unsigned norm16(unsigned v, float a, float b)
{
    const int q = __float2int_rn(__fmaf_rn((float)v, a, b));
    return (unsigned)min(max(q, 0), 65535);
}

constexpr int kHistSlabs = 8;      // row slabs per tile: more CTAs per histogram

__global__ void __launch_bounds__(256)
hist16_kernel(Geom16 g, const unsigned* __restrict__ mm, unsigned* __restrict__ hist)
{
    const int tile = blockIdx.x, slab = blockIdx.y, f = blockIdx.z;
    const int ty = tile / kTiles, tx = tile % kTiles;
    float a, b; norm_coeffs(mm, f, &a, &b);
    unsigned* H = hist + ((size_t)f * kTiles * kTiles + tile) * kHist16;
    const uint16_t* img = g.base + (size_t)f * g.frame_stride_elems;
    const int r0 = (int)((long long)g.th * slab / kHistSlabs), r1 = (int)((long long)g.th * (slab + 1) / kHistSlabs);
    for (int r = r0; r < r1; r++)
    {
        const uint16_t* row = img + (size_t)reflect101(ty * g.th + r, g.h) * g.pitch_elems;
        for (int c = threadIdx.x; c < g.tw; c += blockDim.x)
            atomicAdd(&H[norm16(row[reflect101(tx * g.tw + c, g.w)], a, b)], 1u);
    }
}

constexpr int kLutThreads = 1024, kBinsPerThread = kHist16 / kLutThreads;     // 64

__global__ void __launch_bounds__(kLutThreads)
lut16_kernel(Geom16 g, const unsigned* __restrict__ hist, uint16_t* __restrict__ lut)
{
    __shared__ unsigned s_warp[kLutThreads / 32];
    __shared__ unsigned s_total;
    const int tile = blockIdx.x, f = blockIdx.y, tid = threadIdx.x, lane = tid & 31, wi = tid >> 5;
    const unsigned* H = hist + ((size_t)f * kTiles * kTiles + tile) * kHist16;
    uint16_t* L = lut + ((size_t)f * kTiles * kTiles + tile) * kHist16;
    const int i0 = tid * kBinsPerThread;
    auto block_sum = [&](unsigned v) -> unsigned
    {
        v = __reduce_add_sync(0xffffffffu, v);
        __syncthreads();
        if (lane == 0) s_warp[wi] = v;
        __syncthreads();
        if (tid == 0) { unsigned t = 0; for (int k = 0; k < kLutThreads / 32; k++) t += s_warp[k]; s_total = t; }
        __syncthreads();
        return s_total;
    };
    // clip
    unsigned clipped = 0;
    if (g.clip > 0)
    {
        unsigned mine = 0;
        for (int k = 0; k < kBinsPerThread; k++) { const unsigned h = H[i0 + k]; if (h > (unsigned)g.clip) mine += h - g.clip; }
        clipped = block_sum(mine);
    }
    const unsigned batch = clipped / kHist16, residual = clipped - batch * kHist16;
    const unsigned step = residual ? max(kHist16 / residual, 1u) : 1u;
    // running sums of the clipped, redistributed histogram
    unsigned local = 0;
    for (int k = 0; k < kBinsPerThread; k++)
    {
        const unsigned i = i0 + k;
        unsigned h = H[i];
        if (g.clip > 0)
        {
            h = min(h, (unsigned)g.clip) + batch;
            if (residual && i % step == 0 && i / step < residual) h++;
        }
        local += h;
    }
    // exclusive prefix of `local` over the block
    unsigned incl = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const unsigned v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
    __syncthreads();
    if (lane == 31) s_warp[wi] = incl;
    __syncthreads();
    unsigned before = 0;
    for (int k = 0; k < wi; k++) before += s_warp[k];
    unsigned sum = before + incl - local;
    for (int k = 0; k < kBinsPerThread; k++)
    {
        const unsigned i = i0 + k;
        unsigned h = H[i];
        if (g.clip > 0)
        {
            h = min(h, (unsigned)g.clip) + batch;
            if (residual && i % step == 0 && i / step < residual) h++;
        }
        sum += h;
        const int q = __float2int_rn(__fmul_rn((float)(int)sum, g.lut_scale));
        L[i] = (uint16_t)min(max(q, 0), 65535);
    }
}

__global__ void __launch_bounds__(256)
apply16_kernel(Geom16 g, const unsigned* __restrict__ mm, const uint16_t* __restrict__ lut,
               uint8_t* __restrict__ dst, int dst_pitch, size_t dst_frame_stride)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, f = blockIdx.z;
    if (x >= g.w) return;
    float a, b; norm_coeffs(mm, f, &a, &b);
    const unsigned v = norm16(g.base[(size_t)f * g.frame_stride_elems + (size_t)y * g.pitch_elems + x], a, b);
    const float tyf = __fsub_rn(__fmul_rn((float)y, g.inv_th), 0.5f);
    int ty1 = (int)floorf(tyf), ty2 = ty1 + 1;
    const float ya = __fsub_rn(tyf, (float)ty1), ya1 = __fsub_rn(1.0f, ya);
    ty1 = max(ty1, 0); ty2 = min(ty2, kTiles - 1);
    const float txf = __fsub_rn(__fmul_rn((float)x, g.inv_tw), 0.5f);
    int tx1 = (int)floorf(txf), tx2 = tx1 + 1;
    const float xa = __fsub_rn(txf, (float)tx1), xa1 = __fsub_rn(1.0f, xa);
    tx1 = max(tx1, 0); tx2 = min(tx2, kTiles - 1);
    const uint16_t* L = lut + (size_t)f * kTiles * kTiles * kHist16 + v;
    const float p1 = L[(size_t)(ty1 * kTiles + tx1) * kHist16], p2 = L[(size_t)(ty1 * kTiles + tx2) * kHist16];
    const float q1 = L[(size_t)(ty2 * kTiles + tx1) * kHist16], q2 = L[(size_t)(ty2 * kTiles + tx2) * kHist16];
    const float top = __fadd_rn(__fmul_rn(p1, xa1), __fmul_rn(p2, xa)), bot = __fadd_rn(__fmul_rn(q1, xa1), __fmul_rn(q2, xa));
    const float res = __fadd_rn(__fmul_rn(top, ya1), __fmul_rn(bot, ya));
    const int r16 = min(max(__float2int_rn(res), 0), 65535);
    dst[(size_t)f * dst_frame_stride + (size_t)y * dst_pitch + x] = to8((float)r16);
}
}   // namespace

size_t preproc16_scratch_bytes(int nframes)
{
    return (size_t)nframes * (2 * sizeof(unsigned) + (size_t)kTiles * kTiles * kHist16 * (sizeof(unsigned) + sizeof(uint16_t)));
}

// 16-bit frames -> 8-bit frames as the reference CLI does it: [normalize + CLAHE(clip 8) if clahe], then
// convertTo(CV_8U, 255/65535). scratch: preproc16_scratch_bytes(nframes) (only touched with clahe).
cudaError_t launch_preprocess16(const uint16_t* src, size_t src_frame_stride_elems, int src_pitch_elems, int w, int h, int nframes,
                                bool clahe, uint8_t* dst, int dst_pitch, size_t dst_frame_stride, void* scratch, cudaStream_t stream)
{
    if (w <= 0 || h <= 0 || nframes <= 0) return cudaSuccess;
    if (nframes > 65535 || h > 65535) return cudaErrorInvalidValue;
    Geom16 g;
    g.base = src; g.frame_stride_elems = src_frame_stride_elems; g.pitch_elems = src_pitch_elems; g.w = w; g.h = h; g.n = nframes;
    int we = w, he = h;
    if (w % kTiles || h % kTiles) { we = w + kTiles - w % kTiles; he = h + kTiles - h % kTiles; }
    g.tw = we / kTiles; g.th = he / kTiles;
    const int total = g.tw * g.th;
    g.lut_scale = (float)(kHist16 - 1) / total;
    g.clip = std::max((int)(8.0 * total / kHist16), 1);
    g.inv_tw = 1.0f / g.tw; g.inv_th = 1.0f / g.th;
    const dim3 pgrid((w + 255) / 256, h, nframes);
    if (!clahe)
    {
        convert16to8_kernel<<<pgrid, 256, 0, stream>>>(g, dst, dst_pitch, dst_frame_stride);
        return cudaGetLastError();
    }
    unsigned* mm = (unsigned*)scratch;
    unsigned* hist = mm + 2 * (size_t)nframes;
    uint16_t* lut = (uint16_t*)(hist + (size_t)nframes * kTiles * kTiles * kHist16);
    cudaError_t e = cudaMemsetAsync(mm, 0, sizeof(unsigned) * 2 * nframes + sizeof(unsigned) * (size_t)nframes * kTiles * kTiles * kHist16, stream);
    if (e != cudaSuccess) return e;
    e = cudaMemset2DAsync(mm, 2 * sizeof(unsigned), 0xFF, 2, nframes, stream);      // low 16 bits of every min word = 65535
    if (e != cudaSuccess) return e;
    minmax16_kernel<<<dim3(std::min(h, std::max(1, 4736 / nframes)), nframes), 256, 0, stream>>>(g, mm);
    hist16_kernel<<<dim3(kTiles * kTiles, kHistSlabs, nframes), 256, 0, stream>>>(g, mm, hist);
    lut16_kernel<<<dim3(kTiles * kTiles, nframes), kLutThreads, 0, stream>>>(g, hist, lut);
    apply16_kernel<<<pgrid, 256, 0, stream>>>(g, mm, lut, dst, dst_pitch, dst_frame_stride);
    return cudaGetLastError();
}

}
