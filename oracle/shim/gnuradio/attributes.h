/* Build shim for oracle/_ref: GNU Radio is not installed in this image.  The reference's
 * include/gnuradio/dvbs2rx/api.h only needs these two visibility macros. */
#ifndef ORACLE_SHIM_GR_ATTRIBUTES_H
#define ORACLE_SHIM_GR_ATTRIBUTES_H
#define __GR_ATTR_EXPORT __attribute__((visibility("default")))
#define __GR_ATTR_IMPORT __attribute__((visibility("default")))
#endif
