// The reference CLI's optional --clahe preprocessing, on the GPU:
//     cv::normalize(image, image, 0, 255, NORM_MINMAX);  clahe->apply(image, image1);   // clipLimit 8
// (mrgingham-from-image.cc:43-44, :71-80), which runs before the blur (chess_simple.cu) and the detector.
// OpenCV's arithmetic (third party; pinned to cv2 4.13.0 by tests/test_preproc.py) is reproduced
// exactly: the normalisation is a 256-entry table per frame (scale/shift in double, applied in float
// with one rounding, as OpenCV's FMA build does); CLAHE is 64 tile histograms per frame, clipped and
// redistributed with OpenCV's integer rule, turned into 8-bit tables with a float scale, and applied
// with OpenCV's float bilinear expression (every operation rounded separately, no contraction).
//   P1 minmax_kernel        1 byte/pixel read, per-frame min/max (atomics)
//   P2 norm_lut_kernel      256 threads per frame
//   P3 clahe_lut_kernel     one CTA per (frame, tile): per-warp shared-memory histograms of the normalised
//                           pixels (REFLECT_101 padding when the size is not a multiple of 8), clip,
//                           redistribute, block scan -> table
//   P4 clahe_apply_kernel   1 byte/pixel read + 1 written: four table look-ups and the bilinear blend
#include <cuda_runtime.h>
#include <float.h>
#include <algorithm>

#include "kernels.cuh"

namespace mrgb200
{
namespace
{
constexpr int kTiles = 8, kHist = 256;

// rows are dealt to CTAs; aligned frames are read as 32-bit words (4 pixels per load)
__global__ void __launch_bounds__(256)
minmax_kernel(FrameSet fs, unsigned* __restrict__ mm, int words_ok)
{
    const int f = blockIdx.y;
    const uint8_t* img = fs.base + (size_t)f * fs.frame_stride;
    unsigned mn = 255, mx = 0;
    const int nw = words_ok ? fs.w >> 2 : 0;
    for (int y = blockIdx.x; y < fs.h; y += gridDim.x)
    {
        const uint8_t* row = img + (size_t)y * fs.pitch;
        unsigned lo = 0xffffffffu, hi = 0;
        for (int i = threadIdx.x; i < nw; i += 256)
        {
            const unsigned v = __ldg((const unsigned*)row + i);
            lo = __vminu4(lo, v); hi = __vmaxu4(hi, v);
        }
        if (nw > 0)
        {
#pragma unroll
            for (int k = 0; k < 4; k++) { mn = min(mn, (lo >> (8 * k)) & 255u); mx = max(mx, (hi >> (8 * k)) & 255u); }
        }
        for (int x = 4 * nw + threadIdx.x; x < fs.w; x += 256) { const unsigned v = row[x]; mn = min(mn, v); mx = max(mx, v); }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o)); mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o)); }
    if ((threadIdx.x & 31) == 0) { atomicMin(&mm[2 * f], mn); atomicMax(&mm[2 * f + 1], mx); }
}

__global__ void __launch_bounds__(256)
norm_lut_kernel(const unsigned* __restrict__ mm, uint8_t* __restrict__ nlut)
{
    // cv::normalize NORM_MINMAX to [0,255]: scale = 255 * (1 / (max - min)), 0 when max == min; shift = -min * scale
    const int f = blockIdx.x, v = threadIdx.x;
    const double smin = (double)mm[2 * f], smax = (double)mm[2 * f + 1];
    const double scale = __dmul_rn(255.0, (smax - smin > DBL_EPSILON) ? __ddiv_rn(1.0, __dsub_rn(smax, smin)) : 0.0);
    const double shift = __dsub_rn(0.0, __dmul_rn(smin, scale));
    const float a = __double2float_rn(scale), b = __double2float_rn(shift);
    const int q = __float2int_rn(__fmaf_rn((float)v, a, b));
    nlut[(size_t)f * kHist + v] = (uint8_t)min(max(q, 0), 255);
}

__device__ __forceinline__ int reflect101(int i, int n)
{
    if (n == 1) return 0;
    while (i < 0 || i >= n) i = i < 0 ? -i : 2 * (n - 1) - i;
    return i;
}

struct ClaheGeom { int tw, th, clip; float lut_scale, inv_tw, inv_th; };

__global__ void __launch_bounds__(256)
clahe_lut_kernel(FrameSet fs, ClaheGeom cg, const uint8_t* __restrict__ nlut, uint8_t* __restrict__ tlut)
{
    __shared__ unsigned hw[8][kHist];
    __shared__ int wsum[8];
    __shared__ int s_clipped;
    const int tile = blockIdx.x, f = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tx = tile % kTiles, ty = tile / kTiles;
    const uint8_t* img = fs.base + (size_t)f * fs.frame_stride;
    const uint8_t* nl = nlut ? nlut + (size_t)f * kHist : nullptr;
    for (int i = tid; i < 8 * kHist; i += 256) (&hw[0][0])[i] = 0;
    if (tid == 0) s_clipped = 0;
    __syncthreads();
    // a warp per tile row, lanes along it
    for (int r = warp; r < cg.th; r += 8)
    {
        const uint8_t* row = img + (size_t)reflect101(ty * cg.th + r, fs.h) * fs.pitch;
        for (int c = lane; c < cg.tw; c += 32)
        {
            unsigned v = row[reflect101(tx * cg.tw + c, fs.w)];
            if (nl) v = nl[v];
            atomicAdd(&hw[warp][v], 1u);
        }
    }
    __syncthreads();
    int hcount = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) hcount += (int)hw[k][tid];
    // clip, then spread the excess: the same amount to every bin, the remainder one by one at a fixed stride
    if (cg.clip > 0)
    {
        int excess = max(hcount - cg.clip, 0);
        hcount = min(hcount, cg.clip);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) excess += __shfl_xor_sync(0xffffffffu, excess, o);
        if (lane == 0 && excess) atomicAdd(&s_clipped, excess);
        __syncthreads();
        const int clipped = s_clipped;
        const int batch = clipped / kHist, residual = clipped - batch * kHist;
        hcount += batch;
        if (residual != 0)
        {
            const int step = max(kHist / residual, 1);
            if (tid % step == 0 && tid / step < residual) hcount++;
        }
    }
    // inclusive scan over the 256 bins
    int incl = hcount;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    int base = 0;
    for (int k = 0; k < warp; k++) base += wsum[k];
    const int sum = base + incl;
    const int q = __float2int_rn(__fmul_rn((float)sum, cg.lut_scale));
    tlut[((size_t)f * kTiles * kTiles + tile) * kHist + tid] = (uint8_t)min(max(q, 0), 255);
}

__device__ __forceinline__ unsigned clahe_pixel(const uint8_t* __restrict__ tl, const uint8_t* __restrict__ nl, const ClaheGeom& cg,
                                                int x, unsigned v, int r1, int r2, float ya, float ya1)
{
    const float txf = __fsub_rn(__fmul_rn((float)x, cg.inv_tw), 0.5f);
    int tx1 = (int)floorf(txf), tx2 = tx1 + 1;
    const float xa = __fsub_rn(txf, (float)tx1), xa1 = __fsub_rn(1.0f, xa);
    tx1 = max(tx1, 0); tx2 = min(tx2, kTiles - 1);
    if (nl) v = nl[v];
    const float p1 = tl[(r1 + tx1) * kHist + v], p2 = tl[(r1 + tx2) * kHist + v];
    const float q1 = tl[(r2 + tx1) * kHist + v], q2 = tl[(r2 + tx2) * kHist + v];
    const float res = __fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p1, xa1), __fmul_rn(p2, xa)), ya1),
                                __fmul_rn(__fadd_rn(__fmul_rn(q1, xa1), __fmul_rn(q2, xa)), ya));
    return (unsigned)min(max(__float2int_rn(res), 0), 255);
}

// four pixels per thread; word loads / stores where the frames allow them
__global__ void __launch_bounds__(256)
clahe_apply_kernel(FrameSet fs, ClaheGeom cg, const uint8_t* __restrict__ nlut, const uint8_t* __restrict__ tlut,
                   uint8_t* __restrict__ dst, int dst_pitch, size_t dst_frame_stride, int src_words, int dst_words)
{
    const int f = blockIdx.z, y = blockIdx.y, x0 = (blockIdx.x * 256 + threadIdx.x) * 4;
    if (x0 >= fs.w) return;
    const uint8_t* tl = tlut + (size_t)f * kTiles * kTiles * kHist;
    const uint8_t* nl = nlut ? nlut + (size_t)f * kHist : nullptr;
    const float tyf = __fsub_rn(__fmul_rn((float)y, cg.inv_th), 0.5f);
    int ty1 = (int)floorf(tyf), ty2 = ty1 + 1;
    const float ya = __fsub_rn(tyf, (float)ty1), ya1 = __fsub_rn(1.0f, ya);
    ty1 = max(ty1, 0); ty2 = min(ty2, kTiles - 1);
    const uint8_t* srow = fs.base + (size_t)f * fs.frame_stride + (size_t)y * fs.pitch;
    uint8_t* drow = dst + (size_t)f * dst_frame_stride + (size_t)y * dst_pitch;
    if (x0 + 4 <= fs.w)
    {
        unsigned v;
        if (src_words) v = __ldg((const unsigned*)(srow + x0));
        else v = srow[x0] | (srow[x0 + 1] << 8) | (srow[x0 + 2] << 16) | ((unsigned)srow[x0 + 3] << 24);
        unsigned o = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) o |= clahe_pixel(tl, nl, cg, x0 + k, (v >> (8 * k)) & 255u, ty1 * kTiles, ty2 * kTiles, ya, ya1) << (8 * k);
        if (dst_words) *(unsigned*)(drow + x0) = o;
        else { drow[x0] = (uint8_t)o; drow[x0 + 1] = (uint8_t)(o >> 8); drow[x0 + 2] = (uint8_t)(o >> 16); drow[x0 + 3] = (uint8_t)(o >> 24); }
    }
    else
        for (int x = x0; x < fs.w; x++) drow[x] = (uint8_t)clahe_pixel(tl, nl, cg, x, srow[x], ty1 * kTiles, ty2 * kTiles, ya, ya1);
}
}   // namespace

size_t clahe_scratch_bytes(int nframes) { return (size_t)nframes * (2 * sizeof(unsigned) + kHist + (size_t)kTiles * kTiles * kHist); }

// normalize (optional) + CLAHE(clip_limit, 8x8 tiles) of every frame into dst. scratch: clahe_scratch_bytes(nframes).
cudaError_t launch_normalize_clahe(const FrameSet& fs, bool normalize, double clip_limit, uint8_t* dst, int dst_pitch,
                                   size_t dst_frame_stride, void* scratch, cudaStream_t stream)
{
    if (fs.w <= 0 || fs.h <= 0 || fs.nframes <= 0) return cudaSuccess;
    if (fs.nframes > 65535 || fs.h > 65535) return cudaErrorInvalidValue;
    const int n = fs.nframes;
    unsigned* mm  = (unsigned*)scratch;
    uint8_t* nlut = (uint8_t*)(mm + 2 * n);
    uint8_t* tlut = nlut + (size_t)n * kHist;
    const int src_words = (((uintptr_t)fs.base | (uintptr_t)fs.pitch | fs.frame_stride) & 3) == 0;
    const int dst_words = (((uintptr_t)dst | (uintptr_t)dst_pitch | dst_frame_stride) & 3) == 0;
    if (normalize)
    {
        // min starts at 255, max at 0
        cudaError_t e = cudaMemsetAsync(mm, 0, sizeof(unsigned) * 2 * n, stream);
        if (e != cudaSuccess) return e;
        e = cudaMemset2DAsync(mm, 2 * sizeof(unsigned), 0xFF, 1, n, stream);      // low byte of every min word = 255
        if (e != cudaSuccess) return e;
        const int blocks = std::min(fs.h, std::max(1, 1184 / n));
        minmax_kernel<<<dim3(blocks, n), 256, 0, stream>>>(fs, mm, src_words);
        norm_lut_kernel<<<n, 256, 0, stream>>>(mm, nlut);
    }
    ClaheGeom cg;
    int we = fs.w, he = fs.h;
    if (fs.w % kTiles || fs.h % kTiles) { we = fs.w + kTiles - fs.w % kTiles; he = fs.h + kTiles - fs.h % kTiles; }
    cg.tw = we / kTiles; cg.th = he / kTiles;
    const int total = cg.tw * cg.th;
    cg.lut_scale = (float)(kHist - 1) / total;
    cg.clip = 0;
    if (clip_limit > 0.0) cg.clip = std::max((int)(clip_limit * total / kHist), 1);
    cg.inv_tw = 1.0f / cg.tw; cg.inv_th = 1.0f / cg.th;
    clahe_lut_kernel<<<dim3(kTiles * kTiles, n), 256, 0, stream>>>(fs, cg, normalize ? nlut : nullptr, tlut);
    clahe_apply_kernel<<<dim3((fs.w + 1023) / 1024, fs.h, n), 256, 0, stream>>>(fs, cg, normalize ? nlut : nullptr, tlut,
                                                                                dst, dst_pitch, dst_frame_stride, src_words, dst_words);
    return cudaGetLastError();
}

}
