/* Build shim for oracle/_ref: VOLK is not installed.  lib/pl_descrambler.{h,cc} need volk::vector and the
 * complex multiply kernel; both are restated in their generic form here so that the reference file compiles
 * unmodified (the Gold-sequence construction, lib/pl_descrambler.cc:36-98, is what the pin is about). */
#ifndef ORACLE_SHIM_VOLK_ALLOC_HH
#define ORACLE_SHIM_VOLK_ALLOC_HH
#include <complex>
#include <vector>
namespace volk {
template <class T>
using vector = std::vector<T>;
}
typedef std::complex<float> lv_32fc_t;
inline void volk_32fc_x2_multiply_32fc(lv_32fc_t* c, const lv_32fc_t* a, const lv_32fc_t* b, unsigned int n)
{
    for (unsigned int i = 0; i < n; ++i)
        c[i] = a[i] * b[i];
}
#endif
