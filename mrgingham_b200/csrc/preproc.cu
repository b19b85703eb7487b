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

// rows are dealt to CTAs; aligned frames are read 16 bytes per load
__global__ void __launch_bounds__(256)
minmax_kernel(FrameSet fs, unsigned* __restrict__ mm, int vec_ok)
{
    const int f = blockIdx.y;
    const uint8_t* img = fs.base + (size_t)f * fs.frame_stride;
    unsigned mn = 255, mx = 0;
    const int nv = vec_ok ? fs.w >> 4 : 0;
    unsigned lo = 0xffffffffu, hi = 0;
    for (int y = blockIdx.x; y < fs.h; y += gridDim.x)
    {
        const uint8_t* row = img + (size_t)y * fs.pitch;
#pragma unroll 4
        for (int i = threadIdx.x; i < nv; i += 256)
        {
            const uint4 v = __ldg((const uint4*)row + i);
            lo = __vminu4(__vminu4(lo, v.x), __vminu4(__vminu4(v.y, v.z), v.w));
            hi = __vmaxu4(__vmaxu4(hi, v.x), __vmaxu4(__vmaxu4(v.y, v.z), v.w));
        }
        for (int x = 16 * nv + threadIdx.x; x < fs.w; x += 256) { const unsigned v = row[x]; mn = min(mn, v); mx = max(mx, v); }
    }
#pragma unroll
    for (int k = 0; k < 4; k++) { mn = min(mn, (lo >> (8 * k)) & 255u); mx = max(mx, (hi >> (8 * k)) & 255u); }
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
clahe_lut_kernel(FrameSet fs, ClaheGeom cg, const uint8_t* __restrict__ nlut, uint8_t* __restrict__ tlut, int words_ok)
{
    // histograms of the RAW pixels, two per warp (odd / even lanes) to thin out same-bin collisions; the
    // normalisation table is applied to the 256 bins afterwards instead of to every pixel
    __shared__ unsigned hw[16][kHist];
    __shared__ unsigned hn[kHist];
    __shared__ int wsum[8];
    __shared__ int s_clipped;
    const int tile = blockIdx.x, f = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tx = tile % kTiles, ty = tile / kTiles;
    const uint8_t* img = fs.base + (size_t)f * fs.frame_stride;
    for (int i = tid; i < 16 * kHist; i += 256) (&hw[0][0])[i] = 0;
    hn[tid] = 0;
    if (tid == 0) s_clipped = 0;
    __syncthreads();
    unsigned* h = hw[2 * warp + (lane & 1)];
    const int x0 = tx * cg.tw;
    // 4 pixels per load where the tile's rows are word-aligned and inside the image (no reflection)
    const int nwords = (words_ok && (x0 & 3) == 0 && x0 + cg.tw <= fs.w) ? cg.tw >> 2 : 0;
    for (int r = warp; r < cg.th; r += 8)
    {
        const uint8_t* row = img + (size_t)reflect101(ty * cg.th + r, fs.h) * fs.pitch;
        const unsigned* wrow = (const unsigned*)(row + x0);
        for (int c = lane; c < nwords; c += 32)
        {
            const unsigned v = __ldg(wrow + c);
            const unsigned b0 = v & 255u, b1 = (v >> 8) & 255u, b2 = (v >> 16) & 255u, b3 = v >> 24;
            if (v == b0 * 0x01010101u) atomicAdd(&h[b0], 4u);
            else { atomicAdd(&h[b0], 1u); atomicAdd(&h[b1], 1u); atomicAdd(&h[b2], 1u); atomicAdd(&h[b3], 1u); }
        }
        for (int c = 4 * nwords + lane; c < cg.tw; c += 32)
            atomicAdd(&h[row[reflect101(x0 + c, fs.w)]], 1u);
    }
    __syncthreads();
    {
        unsigned raw = 0;
#pragma unroll
        for (int k = 0; k < 16; k++) raw += hw[k][tid];
        if (nlut) { if (raw) atomicAdd(&hn[nlut[(size_t)f * kHist + tid]], raw); }
        else hn[tid] = raw;
    }
    __syncthreads();
    int hcount = (int)hn[tid];
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
    const int q = min(max(__float2int_rn(__fmul_rn((float)sum, cg.lut_scale)), 0), 255);
    // stored per RAW pixel value: the normalisation folded in, so the apply kernel makes one look-up per table
    __syncthreads();
    hn[tid] = (unsigned)q;
    __syncthreads();
    tlut[((size_t)f * kTiles * kTiles + tile) * kHist + tid] = (uint8_t)hn[nlut ? nlut[(size_t)f * kHist + tid] : tid];
}

// A thread owns 4 adjacent columns and walks kApplyRows rows: the horizontal weights and tile columns are
// computed once. Bytes become floats by planting them in the mantissa of 2^23 (exact), and the blend is
// rounded to nearest-even by adding 1.5 * 2^23 (what rint() does for values in [0, 256)).
constexpr int kApplyRows = 16;

__global__ void __launch_bounds__(256)
clahe_apply_kernel(FrameSet fs, ClaheGeom cg, const uint8_t* __restrict__ tlut,
                   uint8_t* __restrict__ dst, int dst_pitch, size_t dst_frame_stride, int src_words, int dst_words)
{
    const int f = blockIdx.z, x0 = (blockIdx.x * 256 + threadIdx.x) * 4;
    if (x0 >= fs.w) return;
    const uint8_t* tl = tlut + (size_t)f * kTiles * kTiles * kHist;
    float xa[4], xa1[4];
    int c1[4], c2[4];
#pragma unroll
    for (int k = 0; k < 4; k++)
    {
        const float txf = __fsub_rn(__fmul_rn((float)(x0 + k), cg.inv_tw), 0.5f);
        const int tx1 = (int)floorf(txf);
        xa[k] = __fsub_rn(txf, (float)tx1); xa1[k] = __fsub_rn(1.0f, xa[k]);
        c1[k] = max(tx1, 0) * kHist; c2[k] = min(tx1 + 1, kTiles - 1) * kHist;
    }
    const int nx = min(4, fs.w - x0);
    const int yend = min((int)(blockIdx.y + 1) * kApplyRows, fs.h);
    for (int y = blockIdx.y * kApplyRows; y < yend; y++)
    {
        const float tyf = __fsub_rn(__fmul_rn((float)y, cg.inv_th), 0.5f);
        const int ty1 = (int)floorf(tyf);
        const float ya = __fsub_rn(tyf, (float)ty1), ya1 = __fsub_rn(1.0f, ya);
        const uint8_t* t1 = tl + max(ty1, 0) * (kTiles * kHist);
        const uint8_t* t2 = tl + min(ty1 + 1, kTiles - 1) * (kTiles * kHist);
        const uint8_t* srow = fs.base + (size_t)f * fs.frame_stride + (size_t)y * fs.pitch + x0;
        uint8_t* drow = dst + (size_t)f * dst_frame_stride + (size_t)y * dst_pitch + x0;
        unsigned v = 0;
        if (nx == 4 && src_words) v = __ldg((const unsigned*)srow);
        else for (int k = 0; k < nx; k++) v |= (unsigned)srow[k] << (8 * k);
        unsigned o = 0;
#pragma unroll
        for (int k = 0; k < 4; k++)
        {
            const unsigned b = (v >> (8 * k)) & 255u;
            const float p1 = __uint_as_float(0x4B000000u | t1[c1[k] + b]) - 8388608.0f, p2 = __uint_as_float(0x4B000000u | t1[c2[k] + b]) - 8388608.0f;
            const float q1 = __uint_as_float(0x4B000000u | t2[c1[k] + b]) - 8388608.0f, q2 = __uint_as_float(0x4B000000u | t2[c2[k] + b]) - 8388608.0f;
            const float res = __fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p1, xa1[k]), __fmul_rn(p2, xa[k])), ya1),
                                        __fmul_rn(__fadd_rn(__fmul_rn(q1, xa1[k]), __fmul_rn(q2, xa[k])), ya));
            o |= min(__float_as_uint(__fadd_rn(res, 12582912.0f)) & 0x1FFu, 255u) << (8 * k);
        }
        if (nx == 4 && dst_words) *(unsigned*)drow = o;
        else for (int k = 0; k < nx; k++) drow[k] = (uint8_t)(o >> (8 * k));
    }
}

// The same blend for frames whose tiles are at least kApplyRows high (every real image): a CTA's 1024 x 16
// pixel patch then touches at most 2 x 9 interpolation cells, whose four corner tables are interleaved in
// shared memory as one 32-bit word per pixel value, so a pixel costs one shared-memory load instead of
// four global byte gathers.
__global__ void __launch_bounds__(256)
clahe_apply_cells_kernel(FrameSet fs, ClaheGeom cg, const uint8_t* __restrict__ tlut,
                         uint8_t* __restrict__ dst, int dst_pitch, size_t dst_frame_stride, int src_words, int dst_words)
{
    __shared__ unsigned cells[2 * (kTiles + 1) * kHist];
    const int f = blockIdx.z, tid = threadIdx.x, xb = blockIdx.x * 1024, x0 = xb + tid * 4;
    const uint8_t* tl = tlut + (size_t)f * kTiles * kTiles * kHist;
    const int ybeg = blockIdx.y * kApplyRows, yend = min(ybeg + kApplyRows, fs.h);
    // cell index = floor(coordinate / tile - 0.5) + 1, in [0, 8]; cell c blends tiles max(c-1,0) and min(c,7)
    auto cell_of = [](int v, float inv) { return (int)floorf(__fsub_rn(__fmul_rn((float)v, inv), 0.5f)) + 1; };
    const int cx_lo = cell_of(xb, cg.inv_tw), cx_hi = cell_of(min(xb + 1023, fs.w - 1), cg.inv_tw), ncx = cx_hi - cx_lo + 1;
    const int cy_lo = cell_of(ybeg, cg.inv_th), ncy = cell_of(yend - 1, cg.inv_th) - cy_lo + 1;       // 1 or 2
    for (int c = 0; c < ncy * ncx; c++)
    {
        const int cy = cy_lo + c / ncx, cx = cx_lo + c % ncx;
        const int r1 = max(cy - 1, 0) * kTiles, r2 = min(cy, kTiles - 1) * kTiles, c1 = max(cx - 1, 0), c2 = min(cx, kTiles - 1);
        cells[c * kHist + tid] = (unsigned)tl[(r1 + c1) * kHist + tid] | ((unsigned)tl[(r1 + c2) * kHist + tid] << 8) |
                                 ((unsigned)tl[(r2 + c1) * kHist + tid] << 16) | ((unsigned)tl[(r2 + c2) * kHist + tid] << 24);
    }
    __syncthreads();
    if (x0 >= fs.w) return;
    float xa[4], xa1[4];
    int cellx[4];
#pragma unroll
    for (int k = 0; k < 4; k++)
    {
        const float txf = __fsub_rn(__fmul_rn((float)(x0 + k), cg.inv_tw), 0.5f);
        const int tx1 = (int)floorf(txf);
        xa[k] = __fsub_rn(txf, (float)tx1); xa1[k] = __fsub_rn(1.0f, xa[k]);
        cellx[k] = min(tx1 + 1 - cx_lo, ncx - 1) * kHist;       // (the min only guards columns past the image)
    }
    const int nx = min(4, fs.w - x0);
    for (int y = ybeg; y < yend; y++)
    {
        const float tyf = __fsub_rn(__fmul_rn((float)y, cg.inv_th), 0.5f);
        const int ty1 = (int)floorf(tyf);
        const float ya = __fsub_rn(tyf, (float)ty1), ya1 = __fsub_rn(1.0f, ya);
        const unsigned* crow = cells + (ty1 + 1 - cy_lo) * ncx * kHist;
        const uint8_t* srow = fs.base + (size_t)f * fs.frame_stride + (size_t)y * fs.pitch + x0;
        uint8_t* drow = dst + (size_t)f * dst_frame_stride + (size_t)y * dst_pitch + x0;
        unsigned v = 0;
        if (nx == 4 && src_words) v = __ldg((const unsigned*)srow);
        else for (int k = 0; k < nx; k++) v |= (unsigned)srow[k] << (8 * k);
        unsigned o = 0;
#pragma unroll
        for (int k = 0; k < 4; k++)
        {
            const unsigned t = crow[cellx[k] + ((v >> (8 * k)) & 255u)];
            const float p1 = __uint_as_float(__byte_perm(t, 0x4B000000u, 0x7650)) - 8388608.0f;
            const float p2 = __uint_as_float(__byte_perm(t, 0x4B000000u, 0x7651)) - 8388608.0f;
            const float q1 = __uint_as_float(__byte_perm(t, 0x4B000000u, 0x7652)) - 8388608.0f;
            const float q2 = __uint_as_float(__byte_perm(t, 0x4B000000u, 0x7653)) - 8388608.0f;
            const float res = __fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p1, xa1[k]), __fmul_rn(p2, xa[k])), ya1),
                                        __fmul_rn(__fadd_rn(__fmul_rn(q1, xa1[k]), __fmul_rn(q2, xa[k])), ya));
            // a blend of values <= 255 rounds to at most 255: the low byte of the biased sum is the result
            o = __byte_perm(o, __float_as_uint(__fadd_rn(res, 12582912.0f)), k == 0 ? 0x3214 : k == 1 ? 0x3240 : k == 2 ? 0x3410 : 0x4210);
        }
        if (nx == 4 && dst_words) *(unsigned*)drow = o;
        else for (int k = 0; k < nx; k++) drow[k] = (uint8_t)(o >> (8 * k));
    }
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
        const int blocks = std::min(fs.h, std::max(1, 9472 / n));     // 8 waves of resident CTAs: a few rows per CTA
        minmax_kernel<<<dim3(blocks, n), 256, 0, stream>>>(fs, mm, (((uintptr_t)fs.base | (uintptr_t)fs.pitch | fs.frame_stride) & 15) == 0);
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
    clahe_lut_kernel<<<dim3(kTiles * kTiles, n), 256, 0, stream>>>(fs, cg, normalize ? nlut : nullptr, tlut, src_words);
    const dim3 agrid((fs.w + 1023) / 1024, (fs.h + kApplyRows - 1) / kApplyRows, n);
    if (cg.th >= kApplyRows)
        clahe_apply_cells_kernel<<<agrid, 256, 0, stream>>>(fs, cg, tlut, dst, dst_pitch, dst_frame_stride, src_words, dst_words);
    else
        clahe_apply_kernel<<<agrid, 256, 0, stream>>>(fs, cg, tlut, dst, dst_pitch, dst_frame_stride, src_words, dst_words);
    return cudaGetLastError();
}

}
