// Micro-benchmark: issue throughput of the packed-integer / half2 SASS ops the ChESS kernel can be
// built from, alone and in pairs (to learn which ops share an issue pipe on sm_100a).
// Prints warp-instructions per clock per SM for each op / op pair.
#include <cuda_fp16.h>
#include <cstdio>
#include <cstdint>
#include <vector>
#include <string>

#define NCHAIN 8
#define ITERS 2048

enum Op { VABSDIFF4=0, VABSDIFF4ACC, VIADD16X2, VIADDMNMX16, VIMNMX16, IDP4A, PRMT_, IADD3_, LOP3_, HFMA2_, HADD2ABS, IMAD_, HMNMX2_, FFMA_, VIMNMX3_16, SHF_, NOPS };
static const char* opname[] = {"VABSDIFF4","VABSDIFF4.ACC","VIADD.16x2","VIADDMNMX.S16x2","VIMNMX.S16x2","IDP.4A","PRMT","IADD3","LOP3","HFMA2","HADD2|abs|","IMAD","HMNMX2","FFMA","VIMNMX3.S16x2","SHF"};

template<int OP> __device__ __forceinline__ unsigned apply(unsigned x, unsigned y, unsigned z)
{
    if (OP==VABSDIFF4)    return __vabsdiffu4(x,y);
    if (OP==VABSDIFF4ACC) { unsigned d; asm volatile("vabsdiff4.u32.u32.u32.add %0, %1, %2, %3;" : "=r"(d) : "r"(x),"r"(y),"r"(z)); return d; }
    if (OP==VIADD16X2)    return __vadd2(x,y);
    if (OP==VIADDMNMX16)  return __viaddmax_s16x2(x,y,z);
    if (OP==VIMNMX16)     return __vmaxs2(x,y);
    if (OP==IDP4A)        return __dp4a(x,y,z);
    if (OP==PRMT_)        return __byte_perm(x,y,0x5432);
    if (OP==IADD3_)       { unsigned d; asm volatile("add.u32 %0, %1, %2;" : "=r"(d) : "r"(x),"r"(y)); return d; }
    if (OP==LOP3_)        { unsigned d; asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(d) : "r"(x),"r"(y),"r"(z)); return d; }
    if (OP==HFMA2_)       { __half2 a=*(__half2*)&x, b=*(__half2*)&y, c=*(__half2*)&z; __half2 r=__hfma2(a,b,c); return *(unsigned*)&r; }
    if (OP==HADD2ABS)     { unsigned d; asm volatile("{.reg .b32 t; abs.f16x2 t, %1; add.f16x2 %0, t, %2;}" : "=r"(d) : "r"(x),"r"(y)); return d; }
    if (OP==IMAD_)        return x*y+z;
    if (OP==HMNMX2_)      { __half2 a=*(__half2*)&x, b=*(__half2*)&y; __half2 r=__hmax2(a,b); return *(unsigned*)&r; }
    if (OP==FFMA_)        { float r = fmaf(__uint_as_float(x), __uint_as_float(y), __uint_as_float(z)); return __float_as_uint(r); }
    if (OP==VIMNMX3_16)   return __vimax3_s16x2(x,y,z);
    if (OP==SHF_)         return __funnelshift_r(x,y,7);
    return x;
}

template<int OPA, int OPB, int NA, int NB>
__global__ void __launch_bounds__(512) bench(unsigned* out, const unsigned* in, long long* cycles)
{
    unsigned y = in[threadIdx.x & 31], z = in[32 + (threadIdx.x & 31)];
    unsigned xa[NCHAIN], xb[NCHAIN];
#pragma unroll
    for (int j=0;j<NCHAIN;j++){ xa[j]=in[64+j]+threadIdx.x; xb[j]=in[80+j]^threadIdx.x; }
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it=0; it<ITERS; it++) {
#pragma unroll
        for (int r=0;r<NA;r++)
#pragma unroll
            for (int j=0;j<NCHAIN;j++) xa[j]=apply<OPA>(xa[j],y,z);
#pragma unroll
        for (int r=0;r<NB;r++)
#pragma unroll
            for (int j=0;j<NCHAIN;j++) xb[j]=apply<OPB>(xb[j],z,y);
    }
    long long t1 = clock64();
    __syncthreads();
    unsigned acc=0;
#pragma unroll
    for (int j=0;j<NCHAIN;j++) acc ^= xa[j]^xb[j];
    out[blockIdx.x*blockDim.x+threadIdx.x]=acc;
    if (threadIdx.x==0) cycles[blockIdx.x]=t1-t0;
}

template<int OPA, int OPB, int NA, int NB>
static void run(const char* label, unsigned* d_out, unsigned* d_in, long long* d_cyc, int nblk)
{
    bench<OPA,OPB,NA,NB><<<nblk,512>>>(d_out,d_in,d_cyc);
    cudaDeviceSynchronize();
    bench<OPA,OPB,NA,NB><<<nblk,512>>>(d_out,d_in,d_cyc);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<long long> c(nblk);
    cudaMemcpy(c.data(), d_cyc, nblk*sizeof(long long), cudaMemcpyDeviceToHost);
    double avg=0; for (auto v: c) avg+=v; avg/=nblk;
    double ninstr = (double)ITERS*(NA+NB)*NCHAIN*16; // warp-instructions per block (16 warps)
    printf("%-40s %7.3f warp-instr/clk/SM  (%.0f cyc) %s\n", label, ninstr/avg, avg, e==cudaSuccess?"":cudaGetErrorString(e));
}

#define SOLO(OP) { run<OP,OP,1,0>(opname[OP], d_out,d_in,d_cyc,nblk); }
#define PAIR(A,B) { std::string l=std::string(opname[A])+" + "+opname[B]; run<A,B,1,1>(l.c_str(), d_out,d_in,d_cyc,nblk); }

int main()
{
    int nblk = 148;
    unsigned *d_out,*d_in; long long* d_cyc;
    cudaMalloc(&d_out, nblk*512*4); cudaMalloc(&d_in, 4096); cudaMalloc(&d_cyc, nblk*8);
    std::vector<unsigned> h(1024); for (int i=0;i<1024;i++) h[i]=0x3c003c00u ^ (i*2654435761u & 0x03ff03ff);
    cudaMemcpy(d_in,h.data(),4096,cudaMemcpyHostToDevice);
    SOLO(VABSDIFF4) SOLO(VABSDIFF4ACC) SOLO(VIADD16X2) SOLO(VIADDMNMX16) SOLO(VIMNMX16) SOLO(VIMNMX3_16) SOLO(IDP4A) SOLO(PRMT_)
    SOLO(IADD3_) SOLO(LOP3_) SOLO(SHF_) SOLO(HFMA2_) SOLO(HADD2ABS) SOLO(IMAD_) SOLO(HMNMX2_) SOLO(FFMA_)
    PAIR(PRMT_,HFMA2_) PAIR(PRMT_,IMAD_) PAIR(PRMT_,FFMA_) PAIR(PRMT_,IADD3_) PAIR(PRMT_,LOP3_)
    PAIR(VABSDIFF4,HFMA2_) PAIR(VABSDIFF4,PRMT_) PAIR(VABSDIFF4,IMAD_) PAIR(VABSDIFF4,IADD3_)
    PAIR(VIADD16X2,HFMA2_) PAIR(VIADD16X2,PRMT_) PAIR(VIADD16X2,IMAD_) PAIR(VIADD16X2,VABSDIFF4)
    PAIR(VIADDMNMX16,HFMA2_) PAIR(VIADDMNMX16,PRMT_) PAIR(VIADDMNMX16,IMAD_) PAIR(VIMNMX16,PRMT_) PAIR(VIMNMX16,HFMA2_)
    PAIR(IDP4A,PRMT_) PAIR(IDP4A,HFMA2_) PAIR(IDP4A,IMAD_) PAIR(IDP4A,VABSDIFF4)
    PAIR(HADD2ABS,HFMA2_) PAIR(HADD2ABS,PRMT_) PAIR(HADD2ABS,IMAD_) PAIR(HADD2ABS,FFMA_)
    PAIR(HFMA2_,FFMA_) PAIR(HFMA2_,IMAD_) PAIR(IADD3_,IMAD_) PAIR(IADD3_,HFMA2_) PAIR(LOP3_,HFMA2_) PAIR(HMNMX2_,HFMA2_) PAIR(HMNMX2_,PRMT_)
    PAIR(VABSDIFF4ACC,PRMT_) PAIR(VABSDIFF4ACC,HFMA2_) PAIR(VABSDIFF4ACC,IDP4A)
    return 0;
}
