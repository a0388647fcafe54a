/*
 * libsais16.h -- 16-bit-symbol entry points of libsais_cuda (drop-in for the reference's
 * include/libsais16.h: same names, argument meaning, return codes).  HOST pointers; every
 * call computes on the GPU (no CPU fallback: -2 without a usable device).  Symbols are
 * uint16_t; `freq` arrays have 65536 entries.  Semantics are those of the 8-bit functions in
 * libsais.h with the same suffix.  Reference line numbers: include/libsais16.h.
 */
#ifndef LIBSAIS16_CUDA_H
#define LIBSAIS16_CUDA_H 1

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* contexts: one per calling thread, as in the reference */
void * libsais16_create_ctx(void); /* ref :57 */
void * libsais16_create_ctx_omp(int32_t threads); /* ref :66 */
void libsais16_free_ctx(void * ctx); /* ref :73 */
void * libsais16_unbwt_create_ctx(void); /* ref :261 */
void * libsais16_unbwt_create_ctx_omp(int32_t threads); /* ref :270 */
void libsais16_unbwt_free_ctx(void * ctx); /* ref :277 */

/* suffix array / generalized suffix array (separator = symbol 0, T[n-1] must be 0) */
int32_t libsais16(const uint16_t * T, int32_t * SA, int32_t n, int32_t fs, int32_t * freq); /* ref :84 */
int32_t libsais16_gsa(const uint16_t * T, int32_t * SA, int32_t n, int32_t fs, int32_t * freq); /* ref :95 */
int32_t libsais16_int(int32_t * T, int32_t * SA, int32_t n, int32_t k, int32_t fs); /* ref :107 */
int32_t libsais16_ctx(const void * ctx, const uint16_t * T, int32_t * SA, int32_t n, int32_t fs, int32_t * freq); /* ref :119 */
int32_t libsais16_gsa_ctx(const void * ctx, const uint16_t * T, int32_t * SA, int32_t n, int32_t fs, int32_t * freq); /* ref :131 */
int32_t libsais16_omp(const uint16_t * T, int32_t * SA, int32_t n, int32_t fs, int32_t * freq, int32_t threads); /* ref :144 */
int32_t libsais16_gsa_omp(const uint16_t * T, int32_t * SA, int32_t n, int32_t fs, int32_t * freq, int32_t threads); /* ref :156 */
int32_t libsais16_int_omp(int32_t * T, int32_t * SA, int32_t n, int32_t k, int32_t fs, int32_t threads); /* ref :169 */

/* BWT (returns the primary index) and BWT with sampled inverse suffix array (returns 0) */
int32_t libsais16_bwt(const uint16_t * T, uint16_t * U, int32_t * A, int32_t n, int32_t fs, int32_t * freq); /* ref :182 */
int32_t libsais16_bwt_aux(const uint16_t * T, uint16_t * U, int32_t * A, int32_t n, int32_t fs, int32_t * freq, int32_t r, int32_t * I); /* ref :196 */
int32_t libsais16_bwt_ctx(const void * ctx, const uint16_t * T, uint16_t * U, int32_t * A, int32_t n, int32_t fs, int32_t * freq); /* ref :209 */
int32_t libsais16_bwt_aux_ctx(const void * ctx, const uint16_t * T, uint16_t * U, int32_t * A, int32_t n, int32_t fs, int32_t * freq, int32_t r, int32_t * I); /* ref :224 */
int32_t libsais16_bwt_omp(const uint16_t * T, uint16_t * U, int32_t * A, int32_t n, int32_t fs, int32_t * freq, int32_t threads); /* ref :238 */
int32_t libsais16_bwt_aux_omp(const uint16_t * T, uint16_t * U, int32_t * A, int32_t n, int32_t fs, int32_t * freq, int32_t r, int32_t * I, int32_t threads); /* ref :253 */

/* inverse BWT */
int32_t libsais16_unbwt(const uint16_t * T, uint16_t * U, int32_t * A, int32_t n, const int32_t * freq, int32_t i); /* ref :289 */
int32_t libsais16_unbwt_ctx(const void * ctx, const uint16_t * T, uint16_t * U, int32_t * A, int32_t n, const int32_t * freq, int32_t i); /* ref :302 */
int32_t libsais16_unbwt_aux(const uint16_t * T, uint16_t * U, int32_t * A, int32_t n, const int32_t * freq, int32_t r, const int32_t * I); /* ref :315 */
int32_t libsais16_unbwt_aux_ctx(const void * ctx, const uint16_t * T, uint16_t * U, int32_t * A, int32_t n, const int32_t * freq, int32_t r, const int32_t * I); /* ref :329 */
int32_t libsais16_unbwt_omp(const uint16_t * T, uint16_t * U, int32_t * A, int32_t n, const int32_t * freq, int32_t i, int32_t threads); /* ref :343 */
int32_t libsais16_unbwt_aux_omp(const uint16_t * T, uint16_t * U, int32_t * A, int32_t n, const int32_t * freq, int32_t r, const int32_t * I, int32_t threads); /* ref :357 */

/* PLCP from T and SA, PLCP of a generalized SA, LCP from PLCP and SA */
int32_t libsais16_plcp(const uint16_t * T, const int32_t * SA, int32_t * PLCP, int32_t n); /* ref :368 */
int32_t libsais16_plcp_gsa(const uint16_t * T, const int32_t * SA, int32_t * PLCP, int32_t n); /* ref :378 */
int32_t libsais16_lcp(const int32_t * PLCP, const int32_t * SA, int32_t * LCP, int32_t n); /* ref :388 */
int32_t libsais16_plcp_omp(const uint16_t * T, const int32_t * SA, int32_t * PLCP, int32_t n, int32_t threads); /* ref :400 */
int32_t libsais16_plcp_gsa_omp(const uint16_t * T, const int32_t * SA, int32_t * PLCP, int32_t n, int32_t threads); /* ref :411 */
int32_t libsais16_lcp_omp(const int32_t * PLCP, const int32_t * SA, int32_t * LCP, int32_t n, int32_t threads); /* ref :422 */

#ifdef __cplusplus
}
#endif

#endif /* LIBSAIS16_CUDA_H */
