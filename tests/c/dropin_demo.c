/* dropin_demo.c -- a plain C99 caller of the libsais API.  The SAME source is compiled twice by the
 * tests: against the reference (oracle/_ref/libsais_ref.so, built from /root/reference) and against
 * libsais_cuda.so; both binaries must print identical lines.  It exercises what an existing libsais user
 * does: libsais, libsais_bwt (+aux), libsais_unbwt, libsais_plcp, libsais_lcp, libsais_int, libsais64,
 * a context, and the argument-validation / fast-path return codes.
 *
 * usage: dropin_demo <n> <sigma> <seed> [fastpaths]     (deterministic xorshift text) */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "libsais.h"
#include "libsais64.h"

static uint64_t fnv(const void *p, size_t bytes)
{
    const unsigned char *b = (const unsigned char *)p; uint64_t h = 1469598103934665603ull; size_t i;
    for (i = 0; i < bytes; ++i) { h ^= b[i]; h *= 1099511628211ull; }
    return h;
}

int main(int argc, char **argv)
{
    int32_t n = argc > 1 ? atoi(argv[1]) : 6, sigma = argc > 2 ? atoi(argv[2]) : 4, i;
    uint64_t x = argc > 3 ? (uint64_t)atoll(argv[3]) * 2654435761u + 88172645463325252ull : 88172645463325252ull;
    uint8_t *T = (uint8_t *)malloc((size_t)n + 1), *U = (uint8_t *)malloc((size_t)n + 1), *V = (uint8_t *)malloc((size_t)n + 1);
    int32_t *SA = (int32_t *)malloc(sizeof(int32_t) * ((size_t)n + 2)), *A = (int32_t *)malloc(sizeof(int32_t) * ((size_t)n + 2));
    int32_t *P = (int32_t *)malloc(sizeof(int32_t) * ((size_t)n + 2)), *L = (int32_t *)malloc(sizeof(int32_t) * ((size_t)n + 2));
    int32_t *TI = (int32_t *)malloc(sizeof(int32_t) * ((size_t)n + 2)), I[64], freq[256];
    int64_t *SA64 = (int64_t *)malloc(sizeof(int64_t) * ((size_t)n + 2));
    int32_t rc, primary;
    if (argc > 4) {                                    /* fast paths and validation: need no device */
        uint8_t one[1] = { 'x' }; int32_t s1[2] = { -7, -7 };
        printf("null_T %d\n", (int)libsais(NULL, SA, 3, 0, NULL));
        printf("neg_n %d\n", (int)libsais(one, SA, -1, 0, NULL));
        printf("neg_fs %d\n", (int)libsais(one, SA, 1, -1, NULL));
        printf("n0 %d\n", (int)libsais(one, s1, 0, 0, freq));
        printf("n1 %d sa0 %d freqx %d\n", (int)libsais(one, s1, 1, 0, freq), (int)s1[0], (int)freq['x']);
        printf("bwt_n1 %d u0 %c\n", (int)libsais_bwt(one, U, A, 1, 0, NULL), U[0]);
        printf("aux_bad_r %d\n", (int)libsais_bwt_aux(one, U, A, 1, 0, NULL, 3, I));
        printf("unbwt_bad_i %d\n", (int)libsais_unbwt(one, U, A, 1, NULL, 0));
        printf("ctx_null %d\n", (int)libsais_ctx(NULL, one, s1, 1, 0, NULL));
        printf("gsa_no_sep %d\n", (int)libsais_gsa(one, s1, 1, 0, NULL));
        printf("sa64_n1 %d\n", (int)libsais64(one, SA64, 1, 0, NULL));
        return 0;
    }
    for (i = 0; i < n; ++i) { x ^= x << 13; x ^= x >> 7; x ^= x << 17; T[i] = (uint8_t)('a' + (x >> 33) % (uint64_t)sigma); TI[i] = (int32_t)((x >> 20) % 1000); }
    rc = libsais(T, SA, n, 0, freq);
    printf("libsais rc %d sa %016llx freq %016llx\n", (int)rc, (unsigned long long)fnv(SA, sizeof(int32_t) * (size_t)n), (unsigned long long)fnv(freq, sizeof(freq)));
    primary = libsais_bwt(T, U, A, n, 0, NULL);
    printf("bwt primary %d u %016llx\n", (int)primary, (unsigned long long)fnv(U, (size_t)n));
    {
        int32_t r = 2; while ((n - 1) / r + 1 > 64) r *= 2;          /* sampling rate: at most 64 samples */
        rc = libsais_bwt_aux(T, V, A, n, 0, NULL, r, I);
        printf("bwt_aux r %d rc %d I %016llx\n", (int)r, (int)rc, (unsigned long long)fnv(I, sizeof(int32_t) * (size_t)((n - 1) / r + 1)));
        rc = libsais_unbwt_aux(V, U, A, n, NULL, r, I);
        printf("unbwt_aux rc %d same %d\n", (int)rc, (int)(memcmp(U, T, (size_t)n) == 0));
        primary = libsais_bwt(T, U, A, n, 0, NULL);
    }
    rc = libsais_unbwt(U, V, A, n, NULL, primary);
    printf("unbwt rc %d same %d\n", (int)rc, (int)(memcmp(V, T, (size_t)n) == 0));
    rc = libsais_plcp(T, SA, P, n);
    printf("plcp rc %d p %016llx\n", (int)rc, (unsigned long long)fnv(P, sizeof(int32_t) * (size_t)n));
    rc = libsais_lcp(P, SA, L, n);
    printf("lcp rc %d l %016llx\n", (int)rc, (unsigned long long)fnv(L, sizeof(int32_t) * (size_t)n));
    rc = libsais_int(TI, A, n, 1000, 0);
    printf("int rc %d sa %016llx\n", (int)rc, (unsigned long long)fnv(A, sizeof(int32_t) * (size_t)n));
    rc = (int32_t)libsais64(T, SA64, n, 0, NULL);
    printf("sa64 rc %d sa %016llx\n", (int)rc, (unsigned long long)fnv(SA64, sizeof(int64_t) * (size_t)n));
    {
        void *ctx = libsais_create_ctx();
        rc = ctx ? libsais_ctx(ctx, T, A, n, 0, NULL) : -2;
        printf("ctx rc %d same %d\n", (int)rc, (int)(rc == 0 && memcmp(A, SA, sizeof(int32_t) * (size_t)n) == 0));
        libsais_free_ctx(ctx);
    }
    return 0;
}
