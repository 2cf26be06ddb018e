#!/bin/sh
# Diagnostics build of the library with per-phase cycle counters in the LDPC kernel (tools/phase_profile.py):
#   tools/build_prof.sh && DVBS2B200_LIB=$PWD/gr-dvbs2rx_b200/libdvbs2_b200_prof.so python tools/phase_profile.py C1_2 1 1.0 25
set -e
cd "$(dirname "$0")/../gr-dvbs2rx_b200/csrc"
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared -DDVBS2_PHASE_PROFILE \
    capi.cu ldpc_kernel.cu bch_kernel.cu demap_kernel.cu bb_kernel.cu mixed_kernel.cu pl_kernel.cu code_tables.cc -o ../libdvbs2_b200_prof.so
