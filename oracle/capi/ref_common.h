/*
 * oracle/capi/ref_common.h -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Shared prelude for the thin wrappers that compile the UNMODIFIED reference
 * sources, by path, from /root/reference (nothing is copied into this repo).
 * The reference uses the MSVC CRT names _aligned_malloc/_aligned_free
 * (libhog/fhog.h:18,35; trackers/kcf.cpp:160-170,218-226; top/td.cpp:623);
 * map them onto C11 aligned_alloc/free before any reference file is included.
 */
#ifndef ORACLE_REF_COMMON_H
#define ORACLE_REF_COMMON_H

#include <stdlib.h>
#include <string.h>
#include <stdint.h>

#ifndef _aligned_malloc
#define _aligned_malloc(s, a) aligned_alloc((a), ((((size_t)(s)) + (a) - 1) / (a)) * (a))
#endif
#ifndef _aligned_free
#define _aligned_free(p) free(p)
#endif

#define REF_API extern "C" __attribute__((visibility("default")))

#endif
