// Microbenchmark: tcgen05.ld / tcgen05.st throughput when TMEM is used as a per-thread
// spill space (no MMA).  16 warps per CTA, 1 CTA per SM.
#include <cstdio>
#include <cuda_runtime.h>
#include <cstdint>

__device__ __forceinline__ void tm_ld8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
}
__device__ __forceinline__ void tm_st8(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void tm_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tm_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

template <int MODE>   // 0: ld only, 1: st only, 2: ld+st, 3: empty loop (baseline)
__global__ void __launch_bounds__(512, 1) tmem_kernel(int iters, unsigned long long* clk, uint32_t* sink) {
    __shared__ uint32_t tbase;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;"
                     :: "r"((uint32_t)__cvta_generic_to_shared(&tbase)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t taddr = tbase + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)(warp >> 2) * 32;
    uint32_t r[4][8];
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int i = 0; i < 8; ++i) r[q][i] = threadIdx.x * 131 + q * 8 + i;
#pragma unroll
    for (int q = 0; q < 4; ++q) tm_st8(taddr + q * 8, r[q]);
    tm_wait_st();
    __syncthreads();
    unsigned long long t0 = clock64();
    uint32_t accum = 0;
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0 || MODE == 2) {
#pragma unroll
            for (int q = 0; q < 4; ++q) tm_ld8(taddr + q * 8, r[q]);
            tm_wait_ld();
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
                for (int i = 0; i < 8; ++i) accum += r[q][i];
        }
        if (MODE == 1 || MODE == 2) {
#pragma unroll
            for (int q = 0; q < 4; ++q) { r[q][0] += it; tm_st8(taddr + q * 8, r[q]); }
            tm_wait_st();
        }
        if (MODE == 3) accum += it;
    }
    unsigned long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
    // verify content survives: lane-private
    uint32_t chk[8];
    tm_ld8(taddr, chk);
    tm_wait_ld();
    sink[blockIdx.x * blockDim.x + threadIdx.x] = accum + chk[1];
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" :: "r"(tbase));
}

template <int MODE>
void run(const char* name, int iters) {
    unsigned long long* clk; uint32_t* sink;
    cudaMalloc(&clk, 148 * 8); cudaMalloc(&sink, 148 * 512 * 4);
    tmem_kernel<MODE><<<148, 512>>>(10, clk, sink);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    tmem_kernel<MODE><<<148, 512>>>(iters, clk, sink);
    cudaEventRecord(e1);
    cudaError_t err = cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    unsigned long long h[148]; cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
    uint32_t s1; cudaMemcpy(&s1, sink + 1, 4, cudaMemcpyDeviceToHost);
    double cyc = (double)h[0] / iters;
    // per iteration per SM: 16 warps x 32 lanes x 32 regs x 4 B = 64 KiB in each direction
    printf("%-8s err=%s  %.3f ms  %.1f clk/iter/SM  -> %.1f B/clk/SM per direction (64 KiB per iter)  sink=%u\n",
           name, cudaGetErrorString(err), ms, cyc, 65536.0 / cyc, s1);
    cudaFree(clk); cudaFree(sink);
}

int main() {
    run<3>("empty", 20000);
    run<0>("ld", 20000);
    run<1>("st", 20000);
    run<2>("ld+st", 20000);
    return 0;
}
