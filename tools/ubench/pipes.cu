// pipes.cu -- instruction throughput micro-benchmark for the integer/half ops the LDPC kernel
// is built from (run on the B200 box: nvcc -arch=sm_100a -O3 pipes.cu -o pipes && ./pipes).
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <stdint.h>

#define ITERS 4096
#define NACC 8

template <int OP>
__device__ __forceinline__ uint32_t op(uint32_t a, uint32_t b, uint32_t c)
{
    if (OP == 0) return __vmins2(a, b);                 // VIMNMX.S16x2
    if (OP == 1) return __viaddmax_s16x2(a, b, c);           // VIADDMNMX.S16x2
    if (OP == 2) return __vimin3_s16x2(a, b, c);             // VIMNMX3.S16x2
    if (OP == 3) return __vadd2(a, b);                       // VIADD.16x2 ?
    if (OP == 4) return (a & b) ^ c;                         // LOP3
    if (OP == 5) return __byte_perm(a, b, 0x5140);           // PRMT (imm)
    if (OP == 6) return __funnelshift_l(a, b, 7);            // SHF
    if (OP == 7) return a * 33u + b;                         // IMAD
    if (OP == 8) { __half2 r = __hadd2(*(__half2*)&a, *(__half2*)&b); return *(uint32_t*)&r; }
    if (OP == 9) { __half2 r = __hmin2(*(__half2*)&a, *(__half2*)&b); return *(uint32_t*)&r; }
    if (OP == 10) { __half2 r = __hfma2(*(__half2*)&a, *(__half2*)&b, *(__half2*)&c); return *(uint32_t*)&r; }
    if (OP == 11) return a + b + c;                          // IADD3
    if (OP == 12) return (uint32_t)min((int)a, (int)b);      // IMNMX / VIMNMX
    if (OP == 13) return (uint32_t)__viaddmax_s32((int)a, (int)b, (int)c);
    if (OP == 14) return __byte_perm(a, b, c);               // PRMT (reg selector)
    if (OP == 15) return __vsub2(a, b);
    if (OP == 16) return a << (b & 31);                      // SHF / SHL variable
    if (OP == 17) { __half2 r = __hmul2(*(__half2*)&a, *(__half2*)&b); return *(uint32_t*)&r; }
    return a;
}

template <int OP, int OP2>
__global__ void __launch_bounds__(1024) bench(uint32_t* out, uint32_t seed, long long* cycles)
{
    uint32_t acc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i)
        acc[i] = seed * (threadIdx.x + 1 + i);
    uint32_t b = seed ^ 0x01230123u, c = seed + 0x00050003u;
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) {
            acc[i] = op<OP>(acc[i], b, c);
            if (OP2 >= 0)
                acc[i] = op < OP2 < 0 ? 0 : OP2 > (acc[i], c, b);
        }
    }
    long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i)
        s ^= acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0)
        *cycles = t1 - t0;
}

template <int OP, int OP2>
void run(const char* name, uint32_t* d_out, long long* d_cyc)
{
    bench<OP, OP2><<<148, 1024>>>(d_out, 12345u, d_cyc);
    bench<OP, OP2><<<148, 1024>>>(d_out, 12345u, d_cyc);
    cudaDeviceSynchronize();
    long long cyc;
    cudaMemcpy(&cyc, d_cyc, sizeof(cyc), cudaMemcpyDeviceToHost);
    double nops = (double)ITERS * NACC * (OP2 >= 0 ? 2 : 1) * 32; // warp-instructions per SM: 32 warps
    printf("%-34s %8.3f warp-instr/cycle/SM  (%lld cycles)\n", name, nops / (double)cyc, cyc);
}

int main()
{
    uint32_t* d_out;
    long long* d_cyc;
    cudaMalloc(&d_out, 148 * 1024 * 4);
    cudaMalloc(&d_cyc, 8);
    run<0, -1>("VIMNMX.S16x2", d_out, d_cyc);
    run<1, -1>("VIADDMNMX.S16x2", d_out, d_cyc);
    run<2, -1>("VIMNMX3.S16x2", d_out, d_cyc);
    run<3, -1>("vadd2 (VIADD.16x2?)", d_out, d_cyc);
    run<15, -1>("vsub2", d_out, d_cyc);
    run<4, -1>("LOP3", d_out, d_cyc);
    run<5, -1>("PRMT imm", d_out, d_cyc);
    run<14, -1>("PRMT reg", d_out, d_cyc);
    run<6, -1>("SHF", d_out, d_cyc);
    run<16, -1>("SHL var", d_out, d_cyc);
    run<7, -1>("IMAD", d_out, d_cyc);
    run<11, -1>("IADD3", d_out, d_cyc);
    run<12, -1>("IMNMX s32", d_out, d_cyc);
    run<13, -1>("VIADDMNMX s32", d_out, d_cyc);
    run<8, -1>("HADD2", d_out, d_cyc);
    run<17, -1>("HMUL2", d_out, d_cyc);
    run<9, -1>("HMNMX2", d_out, d_cyc);
    run<10, -1>("HFMA2", d_out, d_cyc);
    run<0, 7>("VIMNMX.S16x2 + IMAD", d_out, d_cyc);
    run<0, 4>("VIMNMX.S16x2 + LOP3", d_out, d_cyc);
    run<0, 8>("VIMNMX.S16x2 + HADD2", d_out, d_cyc);
    run<4, 7>("LOP3 + IMAD", d_out, d_cyc);
    run<4, 8>("LOP3 + HADD2", d_out, d_cyc);
    run<9, 8>("HMNMX2 + HADD2", d_out, d_cyc);
    run<9, 4>("HMNMX2 + LOP3", d_out, d_cyc);
    run<5, 7>("PRMT + IMAD", d_out, d_cyc);
    run<1, 10>("VIADDMNMX.S16x2 + HFMA2", d_out, d_cyc);
    run<3, 4>("vadd2 + LOP3", d_out, d_cyc);
    run<3, 7>("vadd2 + IMAD", d_out, d_cyc);
    return 0;
}
