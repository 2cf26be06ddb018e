/* Build shim for oracle/_ref: google/cpu_features is not installed.  x86-64 only (this image). */
#ifndef ORACLE_SHIM_CPU_FEATURES_MACROS_H
#define ORACLE_SHIM_CPU_FEATURES_MACROS_H
#if defined(__x86_64__) || defined(__i386__)
#define CPU_FEATURES_ARCH_X86
#endif
#endif
