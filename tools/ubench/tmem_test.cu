#include <stdint.h>
#include <cstdio>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t tmem_ld(uint32_t taddr) { uint32_t r; asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(taddr)); asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); return r; }
__device__ __forceinline__ void tmem_st(uint32_t taddr, uint32_t v) { asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" :: "r"(taddr), "r"(v) : "memory"); asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__global__ void k(uint32_t* out, int iters)
{
    __shared__ uint32_t s_base;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&s_base)), "n"(128) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = s_base;
    const uint32_t taddr = base + (((uint32_t)(warp & 3) * 32u) << 16) + (uint32_t)(warp >> 2) * 3u + 5u;
    tmem_st(taddr, threadIdx.x * 7u + blockIdx.x);
    uint32_t acc = 0;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        uint32_t v = tmem_ld(taddr);
        tmem_st(taddr, v + 1u);
    }
    long long t1 = clock64();
    acc = tmem_ld(taddr);
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[gridDim.x * blockDim.x] = (uint32_t)((t1 - t0) / iters);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(base), "n"(128) : "memory");
}
int main()
{
    const int grid = 444, threads = 192, iters = 1000;
    uint32_t* d; cudaMalloc(&d, (grid * threads + 1) * 4);
    k<<<grid, threads>>>(d, iters);
    cudaError_t e = cudaDeviceSynchronize();
    printf("sync: %s\n", cudaGetErrorString(e));
    uint32_t* h = new uint32_t[grid * threads + 1];
    cudaMemcpy(h, d, (grid * threads + 1) * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int b = 0; b < grid; ++b) for (int t = 0; t < threads; ++t) if (h[b * threads + t] != (uint32_t)(t * 7 + b + iters)) ++bad;
    printf("bad %d of %d; cycles per ld+st %u\n", bad, grid * threads, h[grid * threads]);
    return 0;
}
