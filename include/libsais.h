/*
 * include/libsais.h -- drop-in C99 interface of libsais_cuda (B200 / sm_100a implementation).
 *
 * Same prototypes, argument meaning and return codes as the reference interface
 * include/libsais.h of IlyaGrebnov/libsais 2.10.4; every entry cites the reference line it replaces.
 * All pointers are HOST pointers.  Work is done on the GPU; there is no CPU fallback:
 * a missing/failed CUDA device makes every computing call return -2.
 * Return codes: 0 (or the primary index for *_bwt) on success, -1 bad arguments,
 * -2 allocation / CUDA failure.  Device-pointer variants live in libsais_cuda.h.
 */

#ifndef LIBSAIS_H
#define LIBSAIS_H 1

#define LIBSAIS_VERSION_MAJOR   2
#define LIBSAIS_VERSION_MINOR   10
#define LIBSAIS_VERSION_PATCH   4
#define LIBSAIS_VERSION_STRING  "2.10.4"

#include <stdint.h>

#if defined(_WIN32) && defined(LIBSAIS_SHARED)
  #if defined(LIBSAIS_EXPORTS)
    #define LIBSAIS_API __declspec(dllexport)
  #else
    #define LIBSAIS_API __declspec(dllimport)
  #endif
#else
  #define LIBSAIS_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

/* Create a context (device workspace + stream bound to one GPU). NULL on failure.  [replaces include/libsais.h:57] */
LIBSAIS_API void *libsais_create_ctx(void);

#if defined(LIBSAIS_OPENMP)
/* Same; `threads` is validated (<0 -> NULL) and otherwise ignored: the GPU is the parallel resource.  [replaces include/libsais.h:66] */
LIBSAIS_API void *libsais_create_ctx_omp(int32_t threads);
#endif

/* Destroy a context; NULL is a no-op.  [replaces include/libsais.h:73] */
LIBSAIS_API void libsais_free_ctx(void * ctx);

/* Suffix array of T[0..n) into SA[0..n); SA[n..n+fs) untouched; freq[256] optional symbol counts. 0 / -1 / -2.  [replaces include/libsais.h:84] */
LIBSAIS_API int32_t libsais(const uint8_t * T, int32_t * SA, int32_t n, int32_t fs, int32_t * freq);

/* Generalized SA of a 0-separated string collection (T[n-1] must be 0). 0 / -1 / -2.  [replaces include/libsais.h:95] */
LIBSAIS_API int32_t libsais_gsa(const uint8_t * T, int32_t * SA, int32_t n, int32_t fs, int32_t * freq);

/* Suffix array of an int32 string with symbols in [0,k). T is left unmodified. 0 / -1 / -2.  [replaces include/libsais.h:107] */
LIBSAIS_API int32_t libsais_int(int32_t * T, int32_t * SA, int32_t n, int32_t k, int32_t fs);

/* libsais() using a caller-owned context.  [replaces include/libsais.h:119] */
LIBSAIS_API int32_t libsais_ctx(const void * ctx, const uint8_t * T, int32_t * SA, int32_t n, int32_t fs, int32_t * freq);

/* libsais_gsa() using a caller-owned context.  [replaces include/libsais.h:131] */
LIBSAIS_API int32_t libsais_gsa_ctx(const void * ctx, const uint8_t * T, int32_t * SA, int32_t n, int32_t fs, int32_t * freq);

#if defined(LIBSAIS_OPENMP)
/* libsais(); threads<0 -> -1, otherwise ignored.  [replaces include/libsais.h:144] */
LIBSAIS_API int32_t libsais_omp(const uint8_t * T, int32_t * SA, int32_t n, int32_t fs, int32_t * freq, int32_t threads);
#endif

#if defined(LIBSAIS_OPENMP)
/* libsais_gsa(); threads<0 -> -1, otherwise ignored.  [replaces include/libsais.h:156] */
LIBSAIS_API int32_t libsais_gsa_omp(const uint8_t * T, int32_t * SA, int32_t n, int32_t fs, int32_t * freq, int32_t threads);
#endif

#if defined(LIBSAIS_OPENMP)
/* libsais_int(); threads<0 -> -1, otherwise ignored.  [replaces include/libsais.h:169] */
LIBSAIS_API int32_t libsais_int_omp(int32_t * T, int32_t * SA, int32_t n, int32_t k, int32_t fs, int32_t threads);
#endif

/* BWT of T into U (U may alias T); A[0..n+fs) is a required but unused temporary. Returns the primary index (>=1 for n>=1), -1 or -2.  [replaces include/libsais.h:182] */
LIBSAIS_API int32_t libsais_bwt(const uint8_t * T, uint8_t * U, int32_t * A, int32_t n, int32_t fs, int32_t * freq);

/* BWT plus auxiliary indexes I[j] = ISA[j*r]+1, j = 0..(n-1)/r; r a power of two >= 2. 0 / -1 / -2.  [replaces include/libsais.h:196] */
LIBSAIS_API int32_t libsais_bwt_aux(const uint8_t * T, uint8_t * U, int32_t * A, int32_t n, int32_t fs, int32_t * freq, int32_t r, int32_t * I);

/* libsais_bwt() using a caller-owned context.  [replaces include/libsais.h:209] */
LIBSAIS_API int32_t libsais_bwt_ctx(const void * ctx, const uint8_t * T, uint8_t * U, int32_t * A, int32_t n, int32_t fs, int32_t * freq);

/* libsais_bwt_aux() using a caller-owned context.  [replaces include/libsais.h:224] */
LIBSAIS_API int32_t libsais_bwt_aux_ctx(const void * ctx, const uint8_t * T, uint8_t * U, int32_t * A, int32_t n, int32_t fs, int32_t * freq, int32_t r, int32_t * I);

#if defined(LIBSAIS_OPENMP)
/* libsais_bwt(); threads<0 -> -1, otherwise ignored.  [replaces include/libsais.h:238] */
LIBSAIS_API int32_t libsais_bwt_omp(const uint8_t * T, uint8_t * U, int32_t * A, int32_t n, int32_t fs, int32_t * freq, int32_t threads);
#endif

#if defined(LIBSAIS_OPENMP)
/* libsais_bwt_aux(); threads<0 -> -1, otherwise ignored.  [replaces include/libsais.h:253] */
LIBSAIS_API int32_t libsais_bwt_aux_omp(const uint8_t * T, uint8_t * U, int32_t * A, int32_t n, int32_t fs, int32_t * freq, int32_t r, int32_t * I, int32_t threads);
#endif

/* Create a context for inverse BWT (same object type as libsais_create_ctx).  [replaces include/libsais.h:261] */
LIBSAIS_API void *libsais_unbwt_create_ctx(void);

#if defined(LIBSAIS_OPENMP)
/* Same; threads<0 -> NULL, otherwise ignored.  [replaces include/libsais.h:270] */
LIBSAIS_API void *libsais_unbwt_create_ctx_omp(int32_t threads);
#endif

/* Destroy an inverse-BWT context; NULL is a no-op.  [replaces include/libsais.h:277] */
LIBSAIS_API void libsais_unbwt_free_ctx(void * ctx);

/* Inverse BWT of T (primary index i) into U (may alias T); A[0..n] required but unused; freq optional (ignored, recomputed on device). 0 / -1 / -2.  [replaces include/libsais.h:289] */
LIBSAIS_API int32_t libsais_unbwt(const uint8_t * T, uint8_t * U, int32_t * A, int32_t n, const int32_t * freq, int32_t i);

/* libsais_unbwt() using a caller-owned context.  [replaces include/libsais.h:302] */
LIBSAIS_API int32_t libsais_unbwt_ctx(const void * ctx, const uint8_t * T, uint8_t * U, int32_t * A, int32_t n, const int32_t * freq, int32_t i);

/* Inverse BWT with auxiliary indexes (r == n or a power of two >= 2; every I[t] in [1,n]). 0 / -1 / -2.  [replaces include/libsais.h:315] */
LIBSAIS_API int32_t libsais_unbwt_aux(const uint8_t * T, uint8_t * U, int32_t * A, int32_t n, const int32_t * freq, int32_t r, const int32_t * I);

/* libsais_unbwt_aux() using a caller-owned context.  [replaces include/libsais.h:329] */
LIBSAIS_API int32_t libsais_unbwt_aux_ctx(const void * ctx, const uint8_t * T, uint8_t * U, int32_t * A, int32_t n, const int32_t * freq, int32_t r, const int32_t * I);

#if defined(LIBSAIS_OPENMP)
/* libsais_unbwt(); threads<0 -> -1, otherwise ignored.  [replaces include/libsais.h:343] */
LIBSAIS_API int32_t libsais_unbwt_omp(const uint8_t * T, uint8_t * U, int32_t * A, int32_t n, const int32_t * freq, int32_t i, int32_t threads);
#endif

#if defined(LIBSAIS_OPENMP)
/* libsais_unbwt_aux(); threads<0 -> -1, otherwise ignored.  [replaces include/libsais.h:357] */
LIBSAIS_API int32_t libsais_unbwt_aux_omp(const uint8_t * T, uint8_t * U, int32_t * A, int32_t n, const int32_t * freq, int32_t r, const int32_t * I, int32_t threads);
#endif

/* Permuted LCP array from T and its suffix array. 0 / -1 / -2.  [replaces include/libsais.h:368] */
LIBSAIS_API int32_t libsais_plcp(const uint8_t * T, const int32_t * SA, int32_t * PLCP, int32_t n);

/* PLCP for a generalized suffix array (matches stop at the 0 separators).  [replaces include/libsais.h:378] */
LIBSAIS_API int32_t libsais_plcp_gsa(const uint8_t * T, const int32_t * SA, int32_t * PLCP, int32_t n);

/* PLCP for an int32 string.  [replaces include/libsais.h:388] */
LIBSAIS_API int32_t libsais_plcp_int(const int32_t * T, const int32_t * SA, int32_t * PLCP, int32_t n);

/* LCP[i] = PLCP[SA[i]]; LCP may alias SA. 0 / -1 / -2.  [replaces include/libsais.h:398] */
LIBSAIS_API int32_t libsais_lcp(const int32_t * PLCP, const int32_t * SA, int32_t * LCP, int32_t n);

#if defined(LIBSAIS_OPENMP)
/* libsais_plcp(); threads<0 -> -1, otherwise ignored.  [replaces include/libsais.h:410] */
LIBSAIS_API int32_t libsais_plcp_omp(const uint8_t * T, const int32_t * SA, int32_t * PLCP, int32_t n, int32_t threads);
#endif

#if defined(LIBSAIS_OPENMP)
/* libsais_plcp_gsa(); threads<0 -> -1, otherwise ignored.  [replaces include/libsais.h:421] */
LIBSAIS_API int32_t libsais_plcp_gsa_omp(const uint8_t * T, const int32_t * SA, int32_t * PLCP, int32_t n, int32_t threads);
#endif

#if defined(LIBSAIS_OPENMP)
/* libsais_plcp_int(); threads<0 -> -1, otherwise ignored.  [replaces include/libsais.h:432] */
LIBSAIS_API int32_t libsais_plcp_int_omp(const int32_t * T, const int32_t * SA, int32_t * PLCP, int32_t n, int32_t threads);
#endif

#if defined(LIBSAIS_OPENMP)
/* libsais_lcp(); threads<0 -> -1, otherwise ignored.  [replaces include/libsais.h:443] */
LIBSAIS_API int32_t libsais_lcp_omp(const int32_t * PLCP, const int32_t * SA, int32_t * LCP, int32_t n, int32_t threads);
#endif

#ifdef __cplusplus
}
#endif

#endif /* LIBSAIS_H */
