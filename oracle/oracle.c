/*
 * oracle.c -- TEST INFRASTRUCTURE ONLY.  Not part of the product library.
 *
 * A plain-C, single-threaded CPU restatement of what the reference (IlyaGrebnov/libsais
 * 2.10.4) computes on the hot path: suffix array, BWT (+primary / aux indexes), inverse
 * BWT, PLCP and LCP, for 32-bit and 64-bit indexes.  It exists so that tests/, bench.py's
 * cpu_baseline leg and __graft_entry__.smoke() can CHECK the CUDA path.  Nothing in the
 * product (libsais_b200/) may link, import or call it.
 *
 * Parity pin: the reference ships no tests or golden vectors, so this
 * restatement is pinned against (a) the known-answer vectors recorded in SURVEY.md §8c
 * (tests/golden/kat.json) and (b) the UNMODIFIED reference compiled from /root/reference
 * into oracle/_ref/libsais_ref.so (oracle/Makefile `ref` target); tests/test_oracle.py runs
 * both comparisons.
 *
 * The SA core is the textbook SA-IS (Nong, Zhang, Chan 2009) the reference is an optimised
 * form of; stages are labelled with the reference regions they correspond to.  This is an
 * independent, unoptimised statement of the algorithm (one generic integer-alphabet routine,
 * explicit sentinel, byte type array) -- none of the reference's code is reproduced.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef int64_t i64;

/* ------------------------------------------------------------------------------------------
 * SA-IS over an integer string s[0..n) whose LAST symbol is a unique smallest sentinel.
 * Stages (reference src/libsais.c):
 *   classify S/L/LMS + bucket counts        ~ count_and_gather_lms_suffixes        :726-837
 *   bucket boundaries                       ~ initialize_buckets_start_and_end     :1398-1427
 *   seed LMS suffixes at bucket ends        ~ radix_sort_lms_suffixes              :1586-1660
 *   induce L then S (LMS-substring order)   ~ induce_partial_order                 :2157-4101
 *   name LMS substrings                     ~ renumber_and_gather_lms_suffixes     :4103-4300
 *   recurse on the reduced string           ~ libsais_main_32s_recursion           :6668-6877
 *   map back + place sorted LMS             ~ reconstruct/place_lms_suffixes       :4548-4776
 *   induce L then S (final order)           ~ induce_final_order                   :4777-6265
 * ---------------------------------------------------------------------------------------- */
#define T_S 1
#define T_L 0
#define IS_LMS(t, i) ((i) > 0 && (t)[(i)] == T_S && (t)[(i) - 1] == T_L)

static void bucket_bounds(const i64 *s, i64 *bkt, i64 n, i64 K, int end)
{
    i64 i, sum = 0;
    for (i = 0; i < K; ++i) bkt[i] = 0;
    for (i = 0; i < n; ++i) bkt[s[i]]++;
    for (i = 0; i < K; ++i) { sum += bkt[i]; bkt[i] = end ? sum : sum - bkt[i]; }
}

static void induce(const i64 *s, const unsigned char *t, i64 *SA, i64 *bkt, i64 n, i64 K)
{
    i64 i, j;
    bucket_bounds(s, bkt, n, K, 0);                      /* L-type: left to right, bucket heads */
    for (i = 0; i < n; ++i) { j = SA[i] - 1; if (SA[i] > 0 && t[j] == T_L) SA[bkt[s[j]]++] = j; }
    bucket_bounds(s, bkt, n, K, 1);                      /* S-type: right to left, bucket tails */
    for (i = n - 1; i >= 0; --i) { j = SA[i] - 1; if (SA[i] > 0 && t[j] == T_S) SA[--bkt[s[j]]] = j; }
}

static int sais_rec(const i64 *s, i64 *SA, i64 n, i64 K)
{
    i64 i, j, n1, name, prev;
    unsigned char *t;
    i64 *bkt, *s1;

    if (n == 1) { SA[0] = 0; return 0; }
    t = (unsigned char *)malloc((size_t)n);
    bkt = (i64 *)malloc((size_t)K * sizeof(i64));
    if (!t || !bkt) { free(t); free(bkt); return -2; }

    t[n - 1] = T_S;
    for (i = n - 2; i >= 0; --i)
        t[i] = (unsigned char)((s[i] < s[i + 1] || (s[i] == s[i + 1] && t[i + 1] == T_S)) ? T_S : T_L);

    /* stage 1: sort LMS substrings by induction from unsorted LMS seeds */
    for (i = 0; i < n; ++i) SA[i] = -1;
    bucket_bounds(s, bkt, n, K, 1);
    for (i = n - 1; i >= 1; --i) if (IS_LMS(t, i)) SA[--bkt[s[i]]] = i;
    induce(s, t, SA, bkt, n, K);

    /* compact the sorted LMS positions to the front, name them */
    n1 = 0;
    for (i = 0; i < n; ++i) if (IS_LMS(t, SA[i])) SA[n1++] = SA[i];
    for (i = n1; i < n; ++i) SA[i] = -1;
    name = 0; prev = -1;
    for (i = 0; i < n1; ++i) {
        i64 pos = SA[i], d; int diff = 0;
        for (d = 0; ; ++d) {
            if (prev < 0 || s[pos + d] != s[prev + d] || t[pos + d] != t[prev + d]) { diff = 1; break; }
            if (d > 0 && (IS_LMS(t, pos + d) || IS_LMS(t, prev + d))) break;
        }
        if (diff) { ++name; prev = pos; }
        SA[n1 + pos / 2] = name - 1;
    }
    for (i = n - 1, j = n - 1; i >= n1; --i) if (SA[i] >= 0) SA[j--] = SA[i];

    /* stage 2: order of the reduced string */
    s1 = SA + n - n1;
    if (name < n1) {
        int rc = sais_rec(s1, SA, n1, name);
        if (rc) { free(t); free(bkt); return rc; }
    } else {
        for (i = 0; i < n1; ++i) SA[s1[i]] = i;
    }

    /* stage 3: map back, seed sorted LMS suffixes, induce the final order */
    for (i = 1, j = 0; i < n; ++i) if (IS_LMS(t, i)) s1[j++] = i;
    for (i = 0; i < n1; ++i) SA[i] = s1[SA[i]];
    for (i = n1; i < n; ++i) SA[i] = -1;
    bucket_bounds(s, bkt, n, K, 1);
    for (i = n1 - 1; i >= 0; --i) { j = SA[i]; SA[i] = -1; SA[--bkt[s[j]]] = j; }
    induce(s, t, SA, bkt, n, K);

    free(t); free(bkt);
    return 0;
}

/* SA of sym[0..n) (get(i) in [0,k)), "a suffix that is a prefix of another sorts first"
 * (reference contract, include/libsais.h:76-84): realised by appending an explicit sentinel. */
static int sais_generic(const void *T, int width, i64 *SA_out, i64 n, i64 k)
{
    i64 i, *s, *SA; int rc;
    if (n == 0) return 0;
    s = (i64 *)malloc((size_t)(n + 1) * sizeof(i64));
    SA = (i64 *)malloc((size_t)(n + 1) * sizeof(i64));
    if (!s || !SA) { free(s); free(SA); return -2; }
    for (i = 0; i < n; ++i) {
        i64 c = width == 1 ? (i64)((const uint8_t *)T)[i]
              : width == 4 ? (i64)((const int32_t *)T)[i] : ((const i64 *)T)[i];
        s[i] = c + 1;
    }
    s[n] = 0;
    rc = sais_rec(s, SA, n + 1, k + 1);
    if (!rc) for (i = 0; i < n; ++i) SA_out[i] = SA[i + 1];
    free(s); free(SA);
    return rc;
}

static i64 max_symbol_plus1(const void *T, int width, i64 n)
{
    i64 i, m = 0;
    for (i = 0; i < n; ++i) {
        i64 c = width == 4 ? (i64)((const int32_t *)T)[i] : ((const i64 *)T)[i];
        if (c + 1 > m) m = c + 1;
    }
    return m;
}

/* ------------------------------------------------------------------------------------------
 * API wrappers, instantiated for int32 and int64 indexes.  Semantics follow SURVEY.md §8a:
 *   libsais          src/libsais.c:7018-7048     libsais_int      :7050-7063
 *   libsais_bwt      :7097-7121                  libsais_bwt_aux  :7123-7146
 *   libsais_unbwt    :8030-8033                  libsais_unbwt_aux:8035-8064
 *   libsais_plcp     :8363-8397 (phi :8116, kasai :8167)   libsais_lcp :8417-8434 (:8311)
 * ---------------------------------------------------------------------------------------- */
#define DEFINE_API(PFX, IDX)                                                                     \
                                                                                                 \
static void PFX##_count(const uint8_t *T, IDX n, IDX *freq)                                      \
{                                                                                                \
    IDX i; if (!freq) return;                                                                    \
    for (i = 0; i < 256; ++i) freq[i] = 0;                                                       \
    for (i = 0; i < n; ++i) freq[T[i]]++;                                                        \
}                                                                                                \
                                                                                                 \
static IDX PFX##_sa(const void *T, int width, IDX *SA, IDX n, i64 k)                             \
{                                                                                                \
    i64 *tmp, i; int rc;                                                                         \
    tmp = (i64 *)malloc((size_t)(n > 0 ? n : 1) * sizeof(i64));                                  \
    if (!tmp) return -2;                                                                         \
    rc = sais_generic(T, width, tmp, (i64)n, k);                                                 \
    if (!rc) for (i = 0; i < (i64)n; ++i) SA[i] = (IDX)tmp[i];                                   \
    free(tmp);                                                                                   \
    return (IDX)rc;                                                                              \
}                                                                                                \
                                                                                                 \
IDX oracle_##PFX(const uint8_t *T, IDX *SA, IDX n, IDX fs, IDX *freq)                            \
{                                                                                                \
    if (!T || !SA || n < 0 || fs < 0) return -1;                                                 \
    PFX##_count(T, n, freq);                                                                     \
    if (n < 2) { if (n == 1) SA[0] = 0; return 0; }                                              \
    return PFX##_sa(T, 1, SA, n, 256);                                                           \
}                                                                                                \
                                                                                                 \
static IDX PFX##_bwt_core(const uint8_t *T, uint8_t *U, IDX *A, IDX n, IDX *freq, IDX r, IDX *I, \
                          IDX *primary)                                                          \
{                                                                                                \
    IDX i, p0 = -1, rc; uint8_t *out;                                                            \
    PFX##_count(T, n, freq);                                                                     \
    rc = PFX##_sa(T, 1, A, n, 256); if (rc) return rc;                                           \
    out = (uint8_t *)malloc((size_t)n); if (!out) return -2;                                     \
    for (i = 0; i < n; ++i) if (A[i] == 0) p0 = i;                                               \
    *primary = p0 + 1;                                     /* primary = ISA[0] + 1 */            \
    out[0] = T[n - 1];                                                                           \
    for (i = 0; i < n; ++i) {                                                                    \
        if (r > 0 && A[i] % r == 0) I[A[i] / r] = i + 1;   /* I[j] = ISA[j*r] + 1 */             \
        if (i == p0) continue;                             /* the "$" row is dropped */          \
        out[i + (i < p0)] = T[A[i] - 1];                                                         \
    }                                                                                            \
    memcpy(U, out, (size_t)n); free(out);                  /* U may alias T */                   \
    return 0;                                                                                    \
}                                                                                                \
                                                                                                 \
IDX oracle_##PFX##_bwt_aux(const uint8_t *T, uint8_t *U, IDX *A, IDX n, IDX fs, IDX *freq,       \
                           IDX r, IDX *I)                                                        \
{                                                                                                \
    IDX primary;                                                                                 \
    if (!T || !U || !A || n < 0 || fs < 0 || r < 2 || (r & (r - 1)) != 0 || !I) return -1;       \
    if (n <= 1) { PFX##_count(T, n, freq); if (n == 1) U[0] = T[0]; I[0] = n; return 0; }        \
    return PFX##_bwt_core(T, U, A, n, freq, r, I, &primary);                                     \
}                                                                                                \
                                                                                                 \
IDX oracle_##PFX##_bwt(const uint8_t *T, uint8_t *U, IDX *A, IDX n, IDX fs, IDX *freq)           \
{                                                                                                \
    IDX primary = 0, rc;                                                                         \
    if (!T || !U || !A || n < 0 || fs < 0) return -1;                                            \
    if (n <= 1) { PFX##_count(T, n, freq); if (n == 1) U[0] = T[0]; return n; }                  \
    rc = PFX##_bwt_core(T, U, A, n, freq, 0, NULL, &primary);                                    \
    return rc ? rc : primary;                                                                    \
}                                                                                                \
                                                                                                 \
IDX oracle_##PFX##_unbwt_aux(const uint8_t *T, uint8_t *U, IDX *A, IDX n, const IDX *freq,       \
                             IDX r, const IDX *I)                                                \
{                                                                                                \
    IDX t, row, k, primary; i64 C[257], c, sum; IDX *LF; uint8_t *out;                           \
    (void)freq;                                                                                  \
    if (!T || !U || !A || n < 0 || (r != n && (r < 2 || (r & (r - 1)) != 0)) || !I) return -1;   \
    if (n <= 1) { if (I[0] != n) return -1; if (n == 1) U[0] = T[0]; return 0; }                 \
    for (t = 0; t <= (n - 1) / r; ++t) if (I[t] <= 0 || I[t] > n) return -1;                     \
    primary = I[0];                                                                              \
    LF = (IDX *)malloc((size_t)(n + 1) * sizeof(IDX)); out = (uint8_t *)malloc((size_t)n);       \
    if (!LF || !out) { free(LF); free(out); return -2; }                                         \
    /* rows 0..n of the sorted rotations of T$; L'[row] = T[row - (row > primary)], $ at primary */ \
    for (c = 0; c < 257; ++c) C[c] = 0;                                                          \
    for (t = 0; t < n; ++t) C[T[t] + 1]++;                                                       \
    for (c = 0, sum = 1; c < 257; ++c) { sum += C[c]; C[c] = sum; }  /* C[x] = 1 + #{sym < x} */ \
    for (row = 0; row <= n; ++row) {                                                             \
        if (row == primary) { LF[row] = 0; continue; }                                           \
        c = T[row - (row > primary)]; LF[row] = (IDX)C[c]++;                                     \
    }                                                                                            \
    for (k = 0, row = 0; k < n; ++k) {                                                           \
        out[n - 1 - k] = T[row - (row > primary)]; row = LF[row];                                \
    }                                                                                            \
    /* aux indexes must agree with the decoded text's sampled ranks; the reference trusts them */ \
    memcpy(U, out, (size_t)n); free(LF); free(out);                                              \
    return 0;                                                                                    \
}                                                                                                \
                                                                                                 \
IDX oracle_##PFX##_unbwt(const uint8_t *T, uint8_t *U, IDX *A, IDX n, const IDX *freq, IDX i)    \
{                                                                                                \
    return oracle_##PFX##_unbwt_aux(T, U, A, n, freq, n, &i);                                    \
}                                                                                                \
                                                                                                 \
static IDX PFX##_plcp_any(const void *T, int width, const IDX *SA, IDX *PLCP, IDX n)             \
{                                                                                                \
    IDX i, l, prev;                                                                              \
    if (!T || !SA || !PLCP || n < 0) return -1;                                                  \
    if (n <= 1) { if (n == 1) PLCP[0] = 0; return 0; }                                           \
    for (i = 0, prev = n; i < n; ++i) { PLCP[SA[i]] = prev; prev = SA[i]; }      /* phi */       \
    for (i = 0, l = 0; i < n; ++i) {                                             /* kasai */     \
        IDX k = PLCP[i];                                                                         \
        if (k == n) { l = 0; }                                                                   \
        else if (width == 1) { const uint8_t *S = (const uint8_t *)T;                            \
            while (i + l < n && k + l < n && S[i + l] == S[k + l]) ++l; }                        \
        else { const int32_t *S = (const int32_t *)T;                                            \
            while (i + l < n && k + l < n && S[i + l] == S[k + l]) ++l; }                        \
        PLCP[i] = l; if (l > 0) --l;                                                             \
    }                                                                                            \
    return 0;                                                                                    \
}                                                                                                \
                                                                                                 \
IDX oracle_##PFX##_plcp(const uint8_t *T, const IDX *SA, IDX *PLCP, IDX n)                       \
{ return PFX##_plcp_any(T, 1, SA, PLCP, n); }                                                    \
                                                                                                 \
IDX oracle_##PFX##_lcp(const IDX *PLCP, const IDX *SA, IDX *LCP, IDX n)                          \
{                                                                                                \
    IDX i;                                                                                       \
    if (!PLCP || !SA || !LCP || n < 0) return -1;                                                \
    for (i = 0; i < n; ++i) LCP[i] = PLCP[SA[i]];                                                \
    return 0;                                                                                    \
}

DEFINE_API(libsais, int32_t)
DEFINE_API(libsais64, int64_t)

/* integer-alphabet entry points: libsais_int (:7050), libsais_plcp_int (:8399), libsais64_long (libsais64.c:7118) */
int32_t oracle_libsais_int(int32_t *T, int32_t *SA, int32_t n, int32_t k, int32_t fs)
{
    i64 kk;
    if (!T || !SA || n < 0 || fs < 0) return -1;
    if (n < 2) { if (n == 1) SA[0] = 0; return 0; }
    kk = max_symbol_plus1(T, 4, n); if (kk < k) kk = k;
    return libsais_sa(T, 4, SA, n, kk);
}

int32_t oracle_libsais_plcp_int(const int32_t *T, const int32_t *SA, int32_t *PLCP, int32_t n)
{ return libsais_plcp_any(T, 4, SA, PLCP, n); }

int64_t oracle_libsais64_long(int64_t *T, int64_t *SA, int64_t n, int64_t k, int64_t fs)
{
    i64 kk;
    if (!T || !SA || n < 0 || fs < 0) return -1;
    if (n < 2) { if (n == 1) SA[0] = 0; return 0; }
    kk = max_symbol_plus1(T, 8, n); if (kk < k) kk = k;
    return libsais64_sa(T, 8, SA, n, kk);
}

/* Generalized suffix array (reference libsais_gsa, src/libsais.c:7033-7048; GSA induction :3022-3053,
 * :5472-5492): every 0 byte is a distinct terminator ordered by position.  Restated by its definition:
 * the plain SA of the text in which the i-th separator is replaced by symbol i and byte c by m + c. */
#define DEFINE_GSA(PFX, IDX)                                                                     \
static i64 *PFX##_gsa_text(const uint8_t *T, IDX n, i64 *k_out)                                  \
{                                                                                                \
    i64 i, m = 0, z = 0, *X = (i64 *)malloc((size_t)(n > 0 ? n : 1) * sizeof(i64));              \
    if (!X) return NULL;                                                                         \
    for (i = 0; i < (i64)n; ++i) m += T[i] == 0;                                                 \
    for (i = 0; i < (i64)n; ++i) X[i] = T[i] == 0 ? z++ : m + T[i];                              \
    *k_out = m + 256;                                                                            \
    return X;                                                                                    \
}                                                                                                \
IDX oracle_##PFX##_gsa(const uint8_t *T, IDX *SA, IDX n, IDX fs, IDX *freq)                      \
{                                                                                                \
    i64 k, *X; IDX rc;                                                                           \
    if (!T || !SA || n < 0 || (n > 0 && T[n - 1] != 0) || fs < 0) return -1;                     \
    PFX##_count(T, n, freq);                                                                     \
    if (n <= 1) { if (n == 1) SA[0] = 0; return 0; }                                             \
    /* the reference rejects empty members (libsais_main_8u :6886-6889, bucket test on symbol 0): \
       verified exhaustively against the compiled reference for n <= 7: T[0] != 0, no "00" */      \
    if (T[0] == 0) return -1;                                                                    \
    { IDX q; for (q = 1; q < n; ++q) if (T[q] == 0 && T[q - 1] == 0) return -1; }                \
    X = PFX##_gsa_text(T, n, &k); if (!X) return -2;                                             \
    rc = PFX##_sa(X, 8, SA, n, k);                                                               \
    free(X);                                                                                     \
    return rc;                                                                                   \
}                                                                                                \
/* PLCP of a GSA (reference compute_plcp_gsa :8215-8238): matches stop at (exclude) separators */ \
IDX oracle_##PFX##_plcp_gsa(const uint8_t *T, const IDX *SA, IDX *PLCP, IDX n)                   \
{                                                                                                \
    IDX i, l, prev;                                                                              \
    if (!T || !SA || !PLCP || n < 0 || (n > 0 && T[n - 1] != 0)) return -1;                      \
    if (n <= 1) { if (n == 1) PLCP[0] = 0; return 0; }                                           \
    for (i = 0, prev = n; i < n; ++i) { PLCP[SA[i]] = prev; prev = SA[i]; }                      \
    for (i = 0, l = 0; i < n; ++i) {                                                             \
        IDX k = PLCP[i];                                                                         \
        if (k == n) l = 0;                                                                       \
        else while (T[i + l] > 0 && T[i + l] == T[k + l]) ++l;                                   \
        PLCP[i] = l; if (l > 0) --l;                                                             \
    }                                                                                            \
    return 0;                                                                                    \
}

DEFINE_GSA(libsais, int32_t)
DEFINE_GSA(libsais64, int64_t)

/* Brute-force definitional checker (second, independent oracle for tiny n): returns the
 * number of adjacent pairs of SA that are out of order or 0 if SA is the sorted suffix order. */
int64_t oracle_check_sa_bruteforce(const uint8_t *T, const int64_t *SA, int64_t n)
{
    i64 i, bad = 0;
    for (i = 1; i < n; ++i) {
        i64 a = SA[i - 1], b = SA[i], la = n - a, lb = n - b, m = la < lb ? la : lb;
        int c = memcmp(T + a, T + b, (size_t)m);
        if (c > 0 || (c == 0 && la >= lb)) ++bad;
    }
    return bad;
}
