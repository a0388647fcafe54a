/*
 * libsais16x64.h -- 16-bit-symbol entry points of libsais_cuda (drop-in for the reference's
 * include/libsais16x64.h: same names, argument meaning, return codes).  HOST pointers; every
 * call computes on the GPU (no CPU fallback: -2 without a usable device).  Symbols are
 * uint16_t; `freq` arrays have 65536 entries.  Semantics are those of the 8-bit functions in
 * libsais64.h with the same suffix.  Reference line numbers: include/libsais16x64.h.
 */
#ifndef LIBSAIS16X64_CUDA_H
#define LIBSAIS16X64_CUDA_H 1

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* suffix array / generalized suffix array (separator = symbol 0, T[n-1] must be 0) */
int64_t libsais16x64(const uint16_t * T, int64_t * SA, int64_t n, int64_t fs, int64_t * freq); /* ref :61 */
int64_t libsais16x64_gsa(const uint16_t * T, int64_t * SA, int64_t n, int64_t fs, int64_t * freq); /* ref :72 */
int64_t libsais16x64_long(int64_t * T, int64_t * SA, int64_t n, int64_t k, int64_t fs); /* ref :84 */
int64_t libsais16x64_omp(const uint16_t * T, int64_t * SA, int64_t n, int64_t fs, int64_t * freq, int64_t threads); /* ref :97 */
int64_t libsais16x64_gsa_omp(const uint16_t * T, int64_t * SA, int64_t n, int64_t fs, int64_t * freq, int64_t threads); /* ref :109 */
int64_t libsais16x64_long_omp(int64_t * T, int64_t * SA, int64_t n, int64_t k, int64_t fs, int64_t threads); /* ref :122 */

/* BWT (returns the primary index) and BWT with sampled inverse suffix array (returns 0) */
int64_t libsais16x64_bwt(const uint16_t * T, uint16_t * U, int64_t * A, int64_t n, int64_t fs, int64_t * freq); /* ref :135 */
int64_t libsais16x64_bwt_aux(const uint16_t * T, uint16_t * U, int64_t * A, int64_t n, int64_t fs, int64_t * freq, int64_t r, int64_t * I); /* ref :149 */
int64_t libsais16x64_bwt_omp(const uint16_t * T, uint16_t * U, int64_t * A, int64_t n, int64_t fs, int64_t * freq, int64_t threads); /* ref :163 */
int64_t libsais16x64_bwt_aux_omp(const uint16_t * T, uint16_t * U, int64_t * A, int64_t n, int64_t fs, int64_t * freq, int64_t r, int64_t * I, int64_t threads); /* ref :178 */

/* inverse BWT */
int64_t libsais16x64_unbwt(const uint16_t * T, uint16_t * U, int64_t * A, int64_t n, const int64_t * freq, int64_t i); /* ref :191 */
int64_t libsais16x64_unbwt_aux(const uint16_t * T, uint16_t * U, int64_t * A, int64_t n, const int64_t * freq, int64_t r, const int64_t * I); /* ref :204 */
int64_t libsais16x64_unbwt_omp(const uint16_t * T, uint16_t * U, int64_t * A, int64_t n, const int64_t * freq, int64_t i, int64_t threads); /* ref :218 */
int64_t libsais16x64_unbwt_aux_omp(const uint16_t * T, uint16_t * U, int64_t * A, int64_t n, const int64_t * freq, int64_t r, const int64_t * I, int64_t threads); /* ref :232 */

/* PLCP from T and SA, PLCP of a generalized SA, LCP from PLCP and SA */
int64_t libsais16x64_plcp(const uint16_t * T, const int64_t * SA, int64_t * PLCP, int64_t n); /* ref :243 */
int64_t libsais16x64_plcp_gsa(const uint16_t * T, const int64_t * SA, int64_t * PLCP, int64_t n); /* ref :253 */
int64_t libsais16x64_lcp(const int64_t * PLCP, const int64_t * SA, int64_t * LCP, int64_t n); /* ref :263 */
int64_t libsais16x64_plcp_omp(const uint16_t * T, const int64_t * SA, int64_t * PLCP, int64_t n, int64_t threads); /* ref :275 */
int64_t libsais16x64_plcp_gsa_omp(const uint16_t * T, const int64_t * SA, int64_t * PLCP, int64_t n, int64_t threads); /* ref :286 */
int64_t libsais16x64_lcp_omp(const int64_t * PLCP, const int64_t * SA, int64_t * LCP, int64_t n, int64_t threads); /* ref :297 */

#ifdef __cplusplus
}
#endif

#endif /* LIBSAIS16X64_CUDA_H */
