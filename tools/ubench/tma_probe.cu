// Probe: which tensor-map shapes does TMA accept for our 272-byte x 10-row box?
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <vector>
#include <cstdlib>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template<int RANK>
__global__ void probe(const __grid_constant__ CUtensorMap tmap_param, const CUtensorMap* tmap_global, uint8_t* out, int bytes, int c0, int c1, int c2, int fence_mode)
{
    const CUtensorMap* tm = tmap_global ? tmap_global : &tmap_param;
    __shared__ __align__(128) uint8_t buf[4096];
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0)
    {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bar)));
        if (fence_mode == 0) asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (fence_mode == 1) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        if (fence_mode == 2) { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
    }
    __syncthreads();
    if (threadIdx.x == 0)
    {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(&bar)), "r"(bytes) : "memory");
        if (RANK == 3)
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                         :: "r"(smem_u32(buf)), "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(&bar)) : "memory");
        else
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                         :: "r"(smem_u32(buf)), "l"(tm), "r"(c0), "r"(c1), "r"(smem_u32(&bar)) : "memory");
    }
    asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D;\nbra W;\nD:\n}\n" :: "r"(smem_u32(&bar)) : "memory");
    for (int i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = buf[i];
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv)
{
    const int mode = argc > 1 ? atoi(argv[1]) : 0; const int promo = argc > 3 ? atoi(argv[3]) : 2; const int fmode = argc > 4 ? atoi(argv[4]) : 0; const int only = argc > 2 ? atoi(argv[2]) : -1; int vi = -1;
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaFree(0);
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    EncodeTiledFn enc = (EncodeTiledFn)p;
    const int W = 640, H = 480, N = 3, pitch = 640;
    std::vector<uint8_t> h((size_t)W*H*N);
    for (size_t i = 0; i < h.size(); i++) h[i] = (uint8_t)(i * 131 + (i >> 9));
    uint8_t *d, *dout; cudaMalloc(&d, h.size()); cudaMalloc(&dout, 4096);
    cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
    struct V { const char* name; CUtensorMapDataType dt; int es; int rank; int boxw; } vs[] = {
        {"u8  rank3 box256", CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, 3, 256},
        {"u16 rank3 box136", CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, 3, 136},
        {"u16 rank3 box128", CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, 3, 128},
        {"u16 rank2 box136", CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, 2, 136},
        {"u32 rank3 box68 ", CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, 3, 68},
        {"u8  rank2 box256", CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, 2, 256},
    };
    CUtensorMap* dmap; cudaMalloc(&dmap, sizeof(CUtensorMap));
    for (auto& v : vs)
    {
        vi++; if (only >= 0 && vi != only) continue;
        CUtensorMap map;
        cuuint64_t dims[3] = { (cuuint64_t)(W / v.es), (cuuint64_t)(v.rank == 3 ? H : H * N), (cuuint64_t)N };
        cuuint64_t strides[2] = { (cuuint64_t)pitch, (cuuint64_t)pitch * H };
        cuuint32_t box[3] = { (cuuint32_t)v.boxw, 10, 1 }, es[3] = { 1, 1, 1 };
        CUresult r = enc(&map, v.dt, v.rank, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                         (CUtensorMapL2promotion)promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("%s: encode=%d ", v.name, (int)r);
        if (r != CUDA_SUCCESS) { printf("\n"); continue; }
        const int bytes = v.boxw * v.es * 10;
        const int xbyte = argc > 5 ? atoi(argv[5]) : 248, y0 = argc > 6 ? atoi(argv[6]) : 100, f = 1;
        cudaMemset(dout, 0xEE, 4096);
        cudaMemcpy(dmap, &map, sizeof(map), cudaMemcpyHostToDevice);
        if (v.rank == 3) probe<3><<<1, 128>>>(map, mode ? dmap : nullptr, dout, bytes, xbyte / v.es, y0, f, fmode);
        else             probe<2><<<1, 128>>>(map, mode ? dmap : nullptr, dout, bytes, xbyte / v.es, y0 + f * H, 0, fmode);
        cudaError_t e = cudaDeviceSynchronize();
        printf("run=%s ", cudaGetErrorString(e));
        if (e == cudaSuccess)
        {
            std::vector<uint8_t> o(bytes); cudaMemcpy(o.data(), dout, bytes, cudaMemcpyDeviceToHost);
            int bad = 0; const int roww = v.boxw * v.es;
            for (int r2 = 0; r2 < 10; r2++) for (int c = 0; c < roww; c++)
            {
                int gx = xbyte + c, gy = y0 + r2;
                uint8_t want = (gx >= 0 && gx < W && gy >= 0 && gy < H) ? h[(size_t)f*W*H + (size_t)gy*pitch + gx] : 0;
                if (o[r2*roww + c] != want) bad++;
            }
            printf("mismatches=%d", bad);
        }
        else { printf("(context poisoned, stopping)\n"); return 1; }
        printf("\n");
    }
    return 0;
}
