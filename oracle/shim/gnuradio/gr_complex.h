/* Build shim for oracle/_ref: lib/psk.hh only needs the two complex typedefs. */
#ifndef ORACLE_SHIM_GR_COMPLEX_H
#define ORACLE_SHIM_GR_COMPLEX_H
#include <complex>
typedef std::complex<float> gr_complex;
typedef std::complex<double> gr_complexd;
#endif
