/*
 * include/libsais64.h -- drop-in C99 interface of libsais_cuda (B200 / sm_100a implementation).
 *
 * Same prototypes, argument meaning and return codes as the reference interface
 * include/libsais64.h of IlyaGrebnov/libsais 2.10.4; every entry cites the reference line it replaces.
 * All pointers are HOST pointers.  Work is done on the GPU; there is no CPU fallback:
 * a missing/failed CUDA device makes every computing call return -2.
 * Return codes: 0 (or the primary index for *_bwt) on success, -1 bad arguments,
 * -2 allocation / CUDA failure.  Device-pointer variants live in libsais_cuda.h.
 */

#ifndef LIBSAIS64_H
#define LIBSAIS64_H 1

#define LIBSAIS64_VERSION_MAJOR   2
#define LIBSAIS64_VERSION_MINOR   10
#define LIBSAIS64_VERSION_PATCH   4
#define LIBSAIS64_VERSION_STRING  "2.10.4"

#include <stdint.h>

#if defined(_WIN32) && defined(LIBSAIS64_SHARED)
  #if defined(LIBSAIS64_EXPORTS)
    #define LIBSAIS64_API __declspec(dllexport)
  #else
    #define LIBSAIS64_API __declspec(dllimport)
  #endif
#else
  #define LIBSAIS64_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

/* Suffix array with 64-bit indexes (inputs of 2 GB and more). 0 / -1 / -2.  [replaces include/libsais64.h:61] */
LIBSAIS64_API int64_t libsais64(const uint8_t * T, int64_t * SA, int64_t n, int64_t fs, int64_t * freq);

/* Generalized SA, 64-bit indexes.  [replaces include/libsais64.h:72] */
LIBSAIS64_API int64_t libsais64_gsa(const uint8_t * T, int64_t * SA, int64_t n, int64_t fs, int64_t * freq);

/* Suffix array of an int64 string with symbols in [0,k). T is left unmodified.  [replaces include/libsais64.h:84] */
LIBSAIS64_API int64_t libsais64_long(int64_t * T, int64_t * SA, int64_t n, int64_t k, int64_t fs);

#if defined(LIBSAIS_OPENMP)
/* libsais64(); threads<0 -> -1, otherwise ignored.  [replaces include/libsais64.h:97] */
LIBSAIS64_API int64_t libsais64_omp(const uint8_t * T, int64_t * SA, int64_t n, int64_t fs, int64_t * freq, int64_t threads);
#endif

#if defined(LIBSAIS_OPENMP)
/* libsais64_gsa(); threads<0 -> -1, otherwise ignored.  [replaces include/libsais64.h:109] */
LIBSAIS64_API int64_t libsais64_gsa_omp(const uint8_t * T, int64_t * SA, int64_t n, int64_t fs, int64_t * freq, int64_t threads);
#endif

#if defined(LIBSAIS_OPENMP)
/* libsais64_long(); threads<0 -> -1, otherwise ignored.  [replaces include/libsais64.h:122] */
LIBSAIS64_API int64_t libsais64_long_omp(int64_t * T, int64_t * SA, int64_t n, int64_t k, int64_t fs, int64_t threads);
#endif

/* BWT, 64-bit indexes; returns the primary index, -1 or -2.  [replaces include/libsais64.h:135] */
LIBSAIS64_API int64_t libsais64_bwt(const uint8_t * T, uint8_t * U, int64_t * A, int64_t n, int64_t fs, int64_t * freq);

/* BWT plus auxiliary indexes, 64-bit.  [replaces include/libsais64.h:149] */
LIBSAIS64_API int64_t libsais64_bwt_aux(const uint8_t * T, uint8_t * U, int64_t * A, int64_t n, int64_t fs, int64_t * freq, int64_t r, int64_t * I);

#if defined(LIBSAIS_OPENMP)
/* libsais64_bwt(); threads<0 -> -1, otherwise ignored.  [replaces include/libsais64.h:163] */
LIBSAIS64_API int64_t libsais64_bwt_omp(const uint8_t * T, uint8_t * U, int64_t * A, int64_t n, int64_t fs, int64_t * freq, int64_t threads);
#endif

#if defined(LIBSAIS_OPENMP)
/* libsais64_bwt_aux(); threads<0 -> -1, otherwise ignored.  [replaces include/libsais64.h:178] */
LIBSAIS64_API int64_t libsais64_bwt_aux_omp(const uint8_t * T, uint8_t * U, int64_t * A, int64_t n, int64_t fs, int64_t * freq, int64_t r, int64_t * I, int64_t threads);
#endif

/* Inverse BWT, 64-bit indexes.  [replaces include/libsais64.h:191] */
LIBSAIS64_API int64_t libsais64_unbwt(const uint8_t * T, uint8_t * U, int64_t * A, int64_t n, const int64_t * freq, int64_t i);

/* Inverse BWT with auxiliary indexes, 64-bit.  [replaces include/libsais64.h:204] */
LIBSAIS64_API int64_t libsais64_unbwt_aux(const uint8_t * T, uint8_t * U, int64_t * A, int64_t n, const int64_t * freq, int64_t r, const int64_t * I);

#if defined(LIBSAIS_OPENMP)
/* libsais64_unbwt(); threads<0 -> -1, otherwise ignored.  [replaces include/libsais64.h:218] */
LIBSAIS64_API int64_t libsais64_unbwt_omp(const uint8_t * T, uint8_t * U, int64_t * A, int64_t n, const int64_t * freq, int64_t i, int64_t threads);
#endif

#if defined(LIBSAIS_OPENMP)
/* libsais64_unbwt_aux(); threads<0 -> -1, otherwise ignored.  [replaces include/libsais64.h:232] */
LIBSAIS64_API int64_t libsais64_unbwt_aux_omp(const uint8_t * T, uint8_t * U, int64_t * A, int64_t n, const int64_t * freq, int64_t r, const int64_t * I, int64_t threads);
#endif

/* PLCP, 64-bit indexes.  [replaces include/libsais64.h:243] */
LIBSAIS64_API int64_t libsais64_plcp(const uint8_t * T, const int64_t * SA, int64_t * PLCP, int64_t n);

/* PLCP for a generalized suffix array, 64-bit.  [replaces include/libsais64.h:253] */
LIBSAIS64_API int64_t libsais64_plcp_gsa(const uint8_t * T, const int64_t * SA, int64_t * PLCP, int64_t n);

/* LCP from PLCP and SA, 64-bit; LCP may alias SA.  [replaces include/libsais64.h:263] */
LIBSAIS64_API int64_t libsais64_lcp(const int64_t * PLCP, const int64_t * SA, int64_t * LCP, int64_t n);

#if defined(LIBSAIS_OPENMP)
/* libsais64_plcp(); threads<0 -> -1, otherwise ignored.  [replaces include/libsais64.h:275] */
LIBSAIS64_API int64_t libsais64_plcp_omp(const uint8_t * T, const int64_t * SA, int64_t * PLCP, int64_t n, int64_t threads);
#endif

#if defined(LIBSAIS_OPENMP)
/* libsais64_plcp_gsa(); threads<0 -> -1, otherwise ignored.  [replaces include/libsais64.h:286] */
LIBSAIS64_API int64_t libsais64_plcp_gsa_omp(const uint8_t * T, const int64_t * SA, int64_t * PLCP, int64_t n, int64_t threads);
#endif

#if defined(LIBSAIS_OPENMP)
/* libsais64_lcp(); threads<0 -> -1, otherwise ignored.  [replaces include/libsais64.h:297] */
LIBSAIS64_API int64_t libsais64_lcp_omp(const int64_t * PLCP, const int64_t * SA, int64_t * LCP, int64_t n, int64_t threads);
#endif

#ifdef __cplusplus
}
#endif

#endif /* LIBSAIS64_H */
