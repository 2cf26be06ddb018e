// pipes2.cu -- which issue pipe do the candidate instructions of the LDPC pair step use on sm_100a?
// Every op is a dependent chain per accumulator (8 accumulators per thread, 32 warps per SM), written in
// inline PTX so that the compiler cannot simplify it; pairs of ops interleaved in one chain show whether two
// instructions share a pipe (rate of the pair = sum of the two costs) or not (rate = the slower one).
// fp16 operands are DENORMALS on purpose (bit pattern = small integer): the question is whether half2
// arithmetic on them runs at full rate.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 pipes2.cu -o pipes2 && ./pipes2
#include <cuda_runtime.h>
#include <cstdio>
#include <stdint.h>

#define ITERS 2048
#define NACC 8

enum {
    VMIN16, VADD16, VADDMAX16, VMIN3_16, LOP, PRMT_R, IMADOP, HADD, HADDSAT, HMIN, HMINXS, HFMA, HFMARELU, HMAX3, SHL, IADD3OP,
    HSUBABS, VABSDIFF, HNEGABS, NOPS
};

template <int OP>
__device__ __forceinline__ uint32_t op(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t r = a;
    if (OP == VMIN16) asm("min.s16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    if (OP == VADD16) asm("add.s16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    if (OP == VADDMAX16) r = __viaddmax_s16x2(a, b, c);
    if (OP == VMIN3_16) r = __vimin3_s16x2(a, b, c);
    if (OP == LOP) asm("lop3.b32 %0, %1, %2, %3, 0x6a;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    if (OP == PRMT_R) asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    if (OP == IMADOP) asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    if (OP == HADD) asm("add.rn.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    if (OP == HADDSAT) asm("add.rn.sat.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    if (OP == HMIN) asm("min.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    if (OP == HMINXS) asm("min.xorsign.abs.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    if (OP == HFMA) asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    if (OP == HFMARELU) asm("fma.rn.relu.f16x2 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    if (OP == HMAX3) asm("{.reg .b32 t; max.f16x2 t, %1, %2; min.f16x2 %0, t, %3;}" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    if (OP == SHL) asm("shl.b32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    if (OP == IADD3OP) r = a + b + c;
    if (OP == HSUBABS) asm("{.reg .b32 t; abs.f16x2 t, %1; sub.rn.f16x2 %0, t, %2;}" : "=r"(r) : "r"(a), "r"(b));
    if (OP == VABSDIFF) r = __vabsdiffs2(a, b);
    if (OP == HNEGABS) asm("{.reg .b32 t; abs.f16x2 t, %1; fma.rn.relu.f16x2 %0, t, %2, %3;}" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}

template <int OP, int OP2>
__global__ void __launch_bounds__(1024) bench(uint32_t* out, uint32_t seed, uint32_t b, uint32_t c, long long* cycles)
{
    uint32_t acc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i)
        acc[i] = seed + (threadIdx.x & 7) + i;
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) {
            acc[i] = op<OP>(acc[i], b, c);
            if (OP2 >= 0)
                acc[i] = op<(OP2 < 0 ? 0 : OP2)>(acc[i], b, c);
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

static uint32_t* d_out;
static long long* d_cyc;

template <int OP, int OP2>
void run(const char* name, uint32_t seed, uint32_t b, uint32_t c, int nsass)
{
    bench<OP, OP2><<<148, 1024>>>(d_out, seed, b, c, d_cyc);
    bench<OP, OP2><<<148, 1024>>>(d_out, seed, b, c, d_cyc);
    cudaDeviceSynchronize();
    long long cyc;
    cudaMemcpy(&cyc, d_cyc, sizeof(cyc), cudaMemcpyDeviceToHost);
    const double nops = (double)ITERS * NACC * 32; // chain links per SM (32 warps)
    printf("%-40s %7.3f cycles per chain link per SMSP-warp-slot  [%d SASS/link] -> %6.3f warp-instr/cycle/SM\n", name,
           (double)cyc * 4 / nops, nsass, nops * nsass / (double)cyc);
}

int main()
{
    cudaMalloc(&d_out, 148 * 1024 * 4);
    cudaMalloc(&d_cyc, 8);
    // fp16 denormal operands: pattern = integer
    const uint32_t dn = 0x00050003u, seed_dn = 0x00400021u, one = 0x3c003c00u, mone = 0xbc00bc00u, zero = 0u;
    run<VMIN16, -1>("VIMNMX.S16x2", 0x01230123u, 0x00330044u, 0, 1);
    run<VADD16, -1>("VIADD.16x2", 0x01230123u, 0x00030004u, 0, 1);
    run<VADDMAX16, -1>("VIADDMNMX.S16x2", 0x01230123u, 0x00030004u, 0x7fff7fffu, 1);
    run<VMIN3_16, -1>("VIMNMX3.S16x2", 0x01230123u, 0x00330044u, 0x00220055u, 1);
    run<LOP, -1>("LOP3", 0x01230123u, 0x00330044u, 0x0f0f0f0fu, 1);
    run<PRMT_R, -1>("PRMT reg", 0x01230123u, 0x00330044u, 0x5140u, 1);
    run<IMADOP, -1>("IMAD", 0x01230123u, 33u, 7u, 1);
    run<SHL, -1>("SHL reg", 0x01230123u, 1u, 0, 1);
    run<IADD3OP, -1>("IADD3", 0x01230123u, 1u, 3u, 1);
    run<VABSDIFF, -1>("vabsdiffs2", 0x01230123u, 0x00030004u, 0, 1);
    run<HADD, -1>("HADD2 denormal + 0", seed_dn, zero, 0, 1);
    run<HADD, -1>("HADD2 normal", one, zero, 0, 1);
    run<HADDSAT, -1>("HADD2.SAT denormal", seed_dn, zero, 0, 1);
    run<HMIN, -1>("HMNMX2 denormal", seed_dn, 0x00410042u, 0, 1);
    run<HMIN, -1>("HMNMX2 normal", one, 0x40004000u, 0, 1);
    run<HMINXS, -1>("HMNMX2.XORSIGN denormal", seed_dn, 0x80410042u, 0, 1);
    run<HFMA, -1>("HFMA2 denormal*1+0", seed_dn, one, zero, 1);
    run<HFMA, -1>("HFMA2 normal", one, one, zero, 1);
    run<HFMARELU, -1>("HFMA2.RELU denormal", seed_dn, one, zero, 1);
    run<HMAX3, -1>("HMNMX2 max then min (clamp)", seed_dn, 0x00010001u, 0x00600060u, 2);
    run<HSUBABS, -1>("abs + HADD2 (modifier folded?)", seed_dn, zero, 0, 1);
    run<HNEGABS, -1>("abs + HFMA2.RELU (folded?)", seed_dn, one, zero, 1);
    // pairs: same pipe -> cost adds; different pipes -> max
    run<VMIN16, LOP>("VIMNMX.S16x2 + LOP3", 0x01230123u, 0x00330044u, 0x0f0f0f0fu, 2);
    run<VMIN16, IMADOP>("VIMNMX.S16x2 + IMAD", 0x01230123u, 0x00330044u, 7u, 2);
    run<VMIN16, VADD16>("VIMNMX.S16x2 + VIADD.16x2", 0x01230123u, 0x00030004u, 0, 2);
    run<VADD16, IMADOP>("VIADD.16x2 + IMAD", 0x01230123u, 3u, 7u, 2);
    run<VADD16, LOP>("VIADD.16x2 + LOP3", 0x01230123u, 0x00330044u, 0x0f0f0f0fu, 2);
    run<VADD16, HADD>("VIADD.16x2 + HADD2", 0x00230023u, zero, 0, 2);
    run<HADD, IMADOP>("HADD2 + IMAD", seed_dn, zero, 0, 2);
    run<HADD, LOP>("HADD2 + LOP3", seed_dn, zero, zero, 2);
    run<HADD, HFMA>("HADD2 + HFMA2", seed_dn, zero, zero, 2);
    run<HMIN, LOP>("HMNMX2 + LOP3", seed_dn, 0x00410042u, zero, 2);
    run<HMIN, VMIN16>("HMNMX2 + VIMNMX.S16x2", seed_dn, 0x00410042u, 0, 2);
    run<HMIN, HADD>("HMNMX2 + HADD2", seed_dn, 0x00410042u, 0, 2);
    run<HMIN, IMADOP>("HMNMX2 + IMAD", seed_dn, 0x00410042u, 0, 2);
    run<PRMT_R, LOP>("PRMT + LOP3", 0x01230123u, 0x00330044u, 0x5140u, 2);
    run<PRMT_R, HADD>("PRMT + HADD2", 0x00230023u, zero, 0x5410u, 2);
    run<VADDMAX16, HFMA>("VIADDMNMX.S16x2 + HFMA2", 0x00230023u, one, zero, 2);
    run<VMIN3_16, HADD>("VIMNMX3.S16x2 + HADD2", 0x00230023u, zero, zero, 2);
    return 0;
}
