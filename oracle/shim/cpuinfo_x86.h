/* Build shim for oracle/_ref: the two feature bits lib/ldpc_decoder_bb_impl.cc:330-345 reads. */
#ifndef ORACLE_SHIM_CPUINFO_X86_H
#define ORACLE_SHIM_CPUINFO_X86_H
namespace cpu_features {
struct X86Features {
    int avx2, sse4_1;
};
struct X86Info {
    X86Features features;
};
inline X86Info GetX86Info()
{
    X86Info i;
    i.features.avx2 = __builtin_cpu_supports("avx2") ? 1 : 0;
    i.features.sse4_1 = __builtin_cpu_supports("sse4.1") ? 1 : 0;
    return i;
}
} // namespace cpu_features
#endif
