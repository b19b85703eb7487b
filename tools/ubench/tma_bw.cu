// How fast can a column-strip walker pull uint8 rows into shared memory on B200?
// Compares TMA boxes of several shapes against cp.async (LDGSTS) for the access pattern of the
// ChESS kernel: CTA = 128 threads walks down a strip of a 3840x2160 frame, ring of stages, no compute
// (each thread reads one word per staged row so the data is actually consumed).
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(b)), "r"(n) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t par)
{
    asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" :: "r"(smem_u32(b)), "r"(par) : "memory");
}
__device__ __forceinline__ void tma3d(void* dst, const CUtensorMap* m, int c0, int c1, int c2, uint64_t* bar)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 :: "r"(smem_u32(dst)), "l"(m), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)) : "memory");
}

constexpr int kMaxStages = 8;

struct P { int W, H, pitch, nframes, boxw, boxh, stages, look, nstrips; size_t fstride; };

__global__ void __launch_bounds__(128) tma_walk(const __grid_constant__ CUtensorMap map, P p, unsigned* sink)
{
    extern __shared__ __align__(128) uint8_t ring[];
    __shared__ __align__(8) uint64_t full[kMaxStages], empty[kMaxStages];
    const int tid = threadIdx.x;
    const int stage_bytes = ((p.boxw * p.boxh + 127) / 128) * 128;
    if (tid == 0) { for (int s = 0; s < p.stages; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 4); } asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    const int strip = blockIdx.x % p.nstrips, f = blockIdx.x / p.nstrips;
    const int x0 = strip * (p.boxw - 32) - 16;
    const int nit = (p.H + p.boxh - 1) / p.boxh;
    auto issue = [&](int it) {
        const int s = it % p.stages;
        if (it >= p.stages) mbar_wait(&empty[s], ((it / p.stages) - 1) & 1);
        mbar_expect(&full[s], p.boxw * p.boxh);
        tma3d(ring + s * stage_bytes, &map, x0 / 2, it * p.boxh, f, &full[s]);
    };
    if (tid == 0) for (int it = 0; it < p.look && it < nit; it++) issue(it);
    unsigned acc = 0;
    for (int it = 0; it < nit; it++)
    {
        const int s = it % p.stages;
        if (tid == 0 && it + p.look < nit) issue(it + p.look);
        mbar_wait(&full[s], (it / p.stages) & 1);
        const uint32_t* w = (const uint32_t*)(ring + s * stage_bytes);
        for (int r = 0; r < p.boxh; r++) acc += w[(r * p.boxw) / 4 + (tid % (p.boxw / 4))];
        __syncwarp();
        if ((tid & 31) == 0) mbar_arrive(&empty[s]);
    }
    if (acc == 0x12345678u) sink[0] = acc;
}

// cp.async variant: every thread copies 16-byte chunks of the stage; zero fill out of bounds
__global__ void __launch_bounds__(128) cpasync_walk(const uint8_t* base, P p, unsigned* sink)
{
    extern __shared__ __align__(128) uint8_t ring[];
    const int tid = threadIdx.x;
    const int stage_bytes = ((p.boxw * p.boxh + 127) / 128) * 128;
    const int strip = blockIdx.x % p.nstrips, f = blockIdx.x / p.nstrips;
    const int x0 = strip * (p.boxw - 32) - 16;
    const int nit = (p.H + p.boxh - 1) / p.boxh;
    const uint8_t* img = base + (size_t)f * p.fstride;
    const int chunks_per_row = p.boxw / 16, nchunks = chunks_per_row * p.boxh;
    auto issue = [&](int it) {
        const int s = it % p.stages;
        for (int c = tid; c < nchunks; c += 128)
        {
            const int r = c / chunks_per_row, cx = c % chunks_per_row;
            const int gy = it * p.boxh + r, gx = x0 + cx * 16;
            int valid = 0;
            if (gy >= 0 && gy < p.H && gx >= 0 && gx < p.W) valid = min(16, p.W - gx);
            const uint8_t* src = valid ? img + (size_t)gy * p.pitch + gx : img;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(smem_u32(ring + s * stage_bytes + r * p.boxw + cx * 16)), "l"(src), "r"(valid) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    for (int it = 0; it < p.look; it++) { if (it < nit) issue(it); else asm volatile("cp.async.commit_group;" ::: "memory"); }
    unsigned acc = 0;
    for (int it = 0; it < nit; it++)
    {
        const int s = it % p.stages;
        __syncthreads();                       // stage (it+look)%stages == stage of it-(stages-look): all done reading it
        if (it + p.look < nit) issue(it + p.look); else asm volatile("cp.async.commit_group;" ::: "memory");
        // wait until the group of iteration `it` has landed: at most `look` newer groups may be pending
        switch (p.look) {
            case 1: asm volatile("cp.async.wait_group 1;" ::: "memory"); break;
            case 2: asm volatile("cp.async.wait_group 2;" ::: "memory"); break;
            case 3: asm volatile("cp.async.wait_group 3;" ::: "memory"); break;
            case 4: asm volatile("cp.async.wait_group 4;" ::: "memory"); break;
            case 5: asm volatile("cp.async.wait_group 5;" ::: "memory"); break;
            default: asm volatile("cp.async.wait_group 6;" ::: "memory"); break;
        }
        __syncthreads();
        const uint32_t* w = (const uint32_t*)(ring + s * stage_bytes);
        for (int r = 0; r < p.boxh; r++) acc += w[(r * p.boxw) / 4 + (tid % (p.boxw / 4))];
    }
    if (acc == 0x12345678u) sink[0] = acc;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main()
{
    cudaFree(0);
    void* fp = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
    EncodeTiledFn enc = (EncodeTiledFn)fp;
    const int W = 3840, H = 2160, N = 256;
    uint8_t* d; unsigned* sink; cudaMalloc(&d, (size_t)W * H * N); cudaMalloc(&sink, 4);
    cudaMemset(d, 1, (size_t)W * H * N);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    struct C { int boxw, boxh, stages, look, promo; } cs[] = {
        {288, 10, 4, 2, 2}, {288, 11, 8, 6, 2}, {288, 11, 8, 6, 0}, {288, 11, 8, 6, 3}, {288, 22, 4, 2, 2}, {288, 44, 4, 2, 2},
        {160, 22, 4, 2, 2}, {544, 11, 4, 2, 2}, {288, 64, 3, 1, 2}, {288, 128, 2, 1, 2},
    };
    for (auto& c : cs)
    {
        P p; p.W = W; p.H = H; p.pitch = W; p.nframes = N; p.boxw = c.boxw; p.boxh = c.boxh; p.stages = c.stages; p.look = c.look;
        p.nstrips = (W + (c.boxw - 32) - 1) / (c.boxw - 32); p.fstride = (size_t)W * H;
        const int stage_bytes = ((c.boxw * c.boxh + 127) / 128) * 128;
        const int smem = stage_bytes * c.stages;
        CUtensorMap map;
        cuuint64_t dims[3] = { (cuuint64_t)W / 2, (cuuint64_t)H, (cuuint64_t)N }, strides[2] = { (cuuint64_t)W, (cuuint64_t)W * H };
        cuuint32_t box[3] = { (cuuint32_t)c.boxw / 2, (cuuint32_t)c.boxh, 1 }, es[3] = { 1, 1, 1 };
        CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_UINT16, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                         (CUtensorMapL2promotion)c.promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); continue; }
        cudaFuncSetAttribute(tma_walk, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        cudaFuncSetAttribute(cpasync_walk, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        const int grid = p.nstrips * N;
        float best[2] = {1e9f, 1e9f};
        for (int v = 0; v < 2; v++)
            for (int rep = 0; rep < 3; rep++)
            {
                cudaEventRecord(e0);
                if (v == 0) tma_walk<<<grid, 128, smem>>>(map, p, sink); else cpasync_walk<<<grid, 128, smem>>>(d, p, sink);
                cudaEventRecord(e1);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
                float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best[v]) best[v] = ms;
            }
        const double gb = (double)W * H * N / 1e9;
        printf("box %3dx%-3d stages %d look %d promo %d smem %6d : TMA %7.1f GB/s   cp.async %7.1f GB/s  (useful bytes)\n",
               c.boxw, c.boxh, c.stages, c.look, c.promo, smem, gb / (best[0] * 1e-3), gb / (best[1] * 1e-3));
    }
    return 0;
}
