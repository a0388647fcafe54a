/*
 * libsais_cuda.h -- extras of the B200 implementation that the reference interface has no
 * counterpart for: explicit GPU selection, device-pointer entry points (so a benchmark can
 * time device work with CUDA events, SURVEY.md §8b "extras needed by the metric"), and the
 * per-kernel-class / per-round statistics the roofline report is built from.
 *
 * The drop-in boundary itself is include/libsais.h and include/libsais64.h.
 */
#ifndef LIBSAIS_CUDA_H
#define LIBSAIS_CUDA_H 1

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LIBSAIS_CUDA_MAX_KERNEL_CLASSES 32

typedef struct libsais_cuda_stats {
    int32_t  n_classes;                                   /* valid entries below */
    int32_t  n_rounds;                                    /* prefix-doubling rounds of the last SA build (incl. round 0) */
    uint64_t total_launches;                              /* kernels launched by the last call */
    uint64_t launches[LIBSAIS_CUDA_MAX_KERNEL_CLASSES];   /* per kernel class */
    double   ms[LIBSAIS_CUDA_MAX_KERNEL_CLASSES];         /* device time per class (profiling on), CUDA events */
    double   bytes[LIBSAIS_CUDA_MAX_KERNEL_CLASSES];      /* ALGORITHMIC bytes per class */
    double   device_ms;                                   /* first kernel -> last kernel of the last call */
    uint64_t workspace_bytes;                             /* current device workspace */
} libsais_cuda_stats;

typedef struct libsais_cuda_round {
    uint64_t h;          /* symbols already sorted when the round started (0 = initial k-mer sort) */
    uint64_t n_active;   /* suffixes sorted in the round */
    uint64_t n_groups;   /* unresolved groups left after the round */
    int32_t  passes;     /* radix digit passes */
    int32_t  key_bits;   /* key bits sorted */
    double   device_ms;  /* with profiling on: device time of the round's kernels (CUDA events) ... */
    double   bytes;      /* ... and the algorithmic bytes they moved */
} libsais_cuda_round;

/* Number of CUDA devices visible (0 when there is no usable GPU). */
int32_t libsais_cuda_device_count(void);

/* Context bound to GPU `device` (-1: $LIBSAIS_CUDA_DEVICE, else the calling thread's current
 * device).  Same object type as libsais_create_ctx(); free with libsais_free_ctx(). */
void *  libsais_cuda_create_ctx(int32_t device);

/* The cudaStream_t all work of this context is issued on. */
void *  libsais_cuda_stream(const void * ctx);

/* Record a CUDA-event pair around every kernel launch (per-class device times in the stats). */
int32_t libsais_cuda_set_profiling(const void * ctx, int32_t on);

/* Statistics of the last call made with this context (ctx == NULL: calling thread's default). */
int32_t libsais_cuda_get_stats(const void * ctx, libsais_cuda_stats * out);
int32_t libsais_cuda_get_round(const void * ctx, int32_t round, libsais_cuda_round * out);
const char * libsais_cuda_kernel_class_name(int32_t kernel_class);
/* Contexts keep their (grow-only) device workspace between calls; this returns it to the system. */
int32_t libsais_cuda_release_workspace(const void * ctx);
/* Debug aid: raw internal device counters (look-back statistics in -DLSC_LOOKBACK_STATS builds). */
int32_t libsais_cuda_debug_scalars(const void * ctx, uint32_t * out, int32_t count);
/* cudaError_t of the last failure seen by the context (0 = none). */
int32_t libsais_cuda_last_error(const void * ctx);

/* ---- device-pointer entry points: every pointer is DEVICE memory on the context's GPU, the
 * call returns after the work completed on the context's stream.  n < 2^32 - 16.
 * Same results as the host functions they mirror.  Return 0 (bwt: primary index) / -1 / -2. */

/* SA[0..n) (uint32 slots) of T; mirrors libsais() [reference include/libsais.h:84]. */
int64_t libsais_cuda_sa_dev(const void * ctx, const uint8_t * d_T, uint32_t * d_SA, int64_t n);
/* BWT of T into d_U (must not alias d_T) and the primary index; mirrors libsais_bwt() [:182]. */
int64_t libsais_cuda_bwt_dev(const void * ctx, const uint8_t * d_T, uint8_t * d_U, int64_t n);
/* PLCP from T and SA; mirrors libsais_plcp() [:368]. */
int64_t libsais_cuda_plcp_dev(const void * ctx, const uint8_t * d_T, const uint32_t * d_SA, uint32_t * d_PLCP, int64_t n);
/* LCP from PLCP and SA (d_LCP must not alias); mirrors libsais_lcp() [:398]. */
int64_t libsais_cuda_lcp_dev(const void * ctx, const uint32_t * d_PLCP, const uint32_t * d_SA, uint32_t * d_LCP, int64_t n);
/* Inverse BWT; mirrors libsais_unbwt() [:289]. */
int64_t libsais_cuda_unbwt_dev(const void * ctx, const uint8_t * d_B, uint8_t * d_U, int64_t n, int64_t primary);

/* ---- batch of independent BWT blocks over one or more GPUs (BASELINE config 4).
 * HOST pointers (pinned memory lets the copies overlap; pageable buffers go through the pinned staging lanes).
 * Block b is processed on devices[b % ndevices] (devices == NULL: the calling thread's current device); every
 * device runs `lanes` host threads (0 = default 3) with one context each -- the reference's
 * one-context-per-thread model [include/libsais.h:53-55] -- so one block's H2D copy, another's kernels and a
 * third's D2H copy overlap.  primary[b] receives what libsais_bwt() returns for block b [include/libsais.h:182];
 * device_ms (nullable) the device time of each block.  Returns 0, -1 (bad arguments) or -2 (a block failed). */
int32_t libsais_cuda_bwt_batch(const uint8_t * const * T, uint8_t * const * U, const int32_t * n, int32_t * primary, float * device_ms,
                               int32_t nblocks, const int32_t * devices, int32_t ndevices, int32_t lanes);
/* Free the pooled contexts of libsais_cuda_bwt_batch (device workspaces, streams). */
void    libsais_cuda_batch_release(void);

/* ---- suffix array of ONE text over several GPUs of a node (BASELINE config 5): distributed prefix doubling with a sample
 * sort, 64-bit positions; exchanges are P2P stores over NVLink issued by the routing kernel itself.  This is what
 * libsais64() [reference include/libsais64.h:61] runs when n exceeds the single-GPU limit (2^32 - 16); $LIBSAIS_CUDA_DIST=G
 * forces it with G GPUs for any n.  T: host text; SA: host int64[n] or NULL (result stays distributed, for timing);
 * freq: nullable int64[256]; devices == NULL: GPUs 0..ndevices-1.  Returns 0, -1 (arguments) or -2. */
typedef struct libsais_cuda_dist_stats {
    int32_t  n_gpus, rounds, key_symbols, key_bits;
    uint64_t slice_max;             /* largest slice of the suffix array held by one GPU */
    uint64_t active_after_round0;   /* unresolved suffixes after the k-mer sort, all GPUs */
    uint64_t exchanged_bytes;       /* bytes written into peer memory (NVLink), all GPUs */
    double   seconds_total;         /* wall time of the call */
    double   seconds_device;        /* from "packed text resident on every GPU" to "all SA slices final" */
    int32_t  verify;                /* in: non-zero runs the distributed checker (ISA[SA[i]] == i and the neighbour-order rule for
                                       every slot) on the device-resident result; out: 1 = proven correct, -1 = violations found */
    int32_t  reserved;
    uint64_t verify_violations;
    double   phase_seconds[8];      /* wall time on rank 0 between barriers: [0] round-0 keys + splitters + fused route, [1] local sort,
                                       [2] rank stage, [3] ISA scatter (route + store), [4] all later rounds */
} libsais_cuda_dist_stats;
int64_t libsais_cuda_sa64_multi(const uint8_t * T, int64_t * SA, int64_t n, int64_t * freq, const int32_t * devices, int32_t ndevices,
                                libsais_cuda_dist_stats * stats);

/* ---- building blocks of the distributed prefix doubling (texts beyond one GPU's working set; orchestrated by
 * libsais_b200/dist.py: one process per GPU, torch.distributed / NCCL all-to-alls between these calls).
 * Device pointers; every call completes before it returns. */

/* Histogram, alphabet map and bit-packing of the (replicated) text; keeps the packed text in the context.
 * Outputs the symbols per round-0 key and the k-mer width in bits. */
int64_t libsais_cuda_dist_prepare(const void * ctx, const uint8_t * d_T, int64_t n, int32_t * k_out, int32_t * key_bits_out);
/* Round-0 keys of positions [lo, lo+count): (k-mer << 7) | length field (127 = full length), and the positions. */
int64_t libsais_cuda_dist_keys(const void * ctx, int64_t lo, int64_t count, uint64_t * d_keys, uint32_t * d_pos);
/* Onesweep sort of (u64 key, u32 value) / (u32, u32) pairs on key bits [lo_bit, hi_bit); returns 0 when the
 * result is in (d_keys, d_vals), 1 when in the alternate buffers, < 0 on error. */
int64_t libsais_cuda_sort_pairs_dev(const void * ctx, uint64_t * d_keys, uint32_t * d_vals, uint64_t * d_keys_alt, uint32_t * d_vals_alt,
                                    int64_t count, int32_t lo_bit, int32_t hi_bit);
int64_t libsais_cuda_sort_u32_pairs_dev(const void * ctx, uint32_t * d_keys, uint32_t * d_vals, uint32_t * d_keys_alt, uint32_t * d_vals_alt,
                                        int64_t count, int32_t lo_bit, int32_t hi_bit);
/* Rank stage on a sorted slice whose element j sits in global slot slot_base + j (d_slot_in == NULL) or
 * d_slot_in[j]: (position, rank) pairs for all elements, the SA slice, the compacted unresolved suffixes;
 * counts_out[0..1] = unresolved suffixes, unresolved groups. */
int64_t libsais_cuda_rank_stage_dev(const void * ctx, const uint64_t * d_keys, const uint32_t * d_pos, const uint32_t * d_slot_in,
                                    int64_t count, uint32_t slot_base, uint32_t * d_sa_local, uint32_t * d_pair_pos, uint32_t * d_pair_rank,
                                    uint32_t * d_act_pos, uint32_t * d_act_slot, uint32_t * d_act_grp, uint64_t * counts_out);
/* Routing for the all-to-alls.  route: items (a[i], b[i]) are grouped by destination rank min(world-1, (a[i]+add)/block)
 * (items with a[i]+add >= limit are dropped behind the last rank) with one onesweep digit pass; outputs hold a[i]+add and
 * b[i] in destination order, counts_out[0..world) the items per destination.  partition: round-0 sample sort, destination
 * = number of splitters <= key (kept in the key's top byte), counts_out[0..nsplit]. */
int64_t libsais_cuda_dist_route_dev(const void * ctx, const uint32_t * d_a, const uint32_t * d_b, int64_t count, int64_t add, int64_t limit,
                                    int64_t block, int32_t world, uint32_t * d_a_out, uint32_t * d_b_out, uint64_t * counts_out);
int64_t libsais_cuda_dist_partition_dev(const void * ctx, uint64_t * d_keys, uint32_t * d_pos, int64_t count, const uint64_t * d_splitters,
                                        int32_t nsplit, uint64_t * d_keys_out, uint32_t * d_pos_out, uint64_t * counts_out);
/* out[i] = src[idx[i] - idx_offset] (0 when out of range);  dst[idx[i] - idx_offset] = val[i]. */
int64_t libsais_cuda_gather_u32_dev(const void * ctx, const uint32_t * d_src, int64_t src_len, const uint32_t * d_idx, int64_t count,
                                    uint32_t idx_offset, uint32_t * d_out);
int64_t libsais_cuda_scatter_u32_dev(const void * ctx, uint32_t * d_dst, int64_t dst_len, const uint32_t * d_idx, const uint32_t * d_val,
                                     int64_t count, uint32_t idx_offset);

#ifdef __cplusplus
}
#endif

#endif /* LIBSAIS_CUDA_H */
