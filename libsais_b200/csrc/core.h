// core.h -- internal interface between the C API layer (api.cu) and the device pipelines.
#pragma once
#include "ctx.h"

namespace lsc {

// d_scalars / h_scalars slot map
enum ScalarSlot {
    S_FREQ    = 0,     // [0,256): byte histogram
    S_ERR     = 256,   // look-back watchdog / internal error flag (u32 view at the low half)
    S_NACT    = 257,   // active suffixes after the last rank kernel
    S_NGRP    = 258,   // non-singleton groups after the last rank kernel
    S_MAXSYM  = 259,   // max symbol of an integer text
    S_PRIMARY = 260,   // slot of suffix 0 (+1)
    S_BIGGRP  = 261,   // set when an unresolved group is larger than the local-sort tile
    S_TICKET  = 264,   // [264, 272): tickets of the rank kernels (u32 views)
    S_MISC    = 272,   // [272, 304): LUT staging; [312, 320): debug counters
    S_MSD     = 304,   // [304, 308): round-0 MSD path: largest 16-bit bucket, most tiles of a top-level bucket, overflow flag
    S_GSA_TOTAL = 320, // number of separators of a GSA text
    S_GSA_INVALID = 321, // set when the collection has an empty member (reference returns -1)
    S_ROUTE   = 384,   // [384, 449): per-destination counts of the distributed routing
    S_SGRAM   = 512    // [512, 768): s-gram histogram of the text
};

struct SAOptions {
    bool want_sa = true;          // materialise SA (false: BWT-only callers never need it)
    u32 *sa_out = nullptr;        // caller's device buffer for SA instead of an arena array
    u8 *bwt_rows = nullptr;       // byte texts only: rows[slot] = byte preceding the suffix in that SA slot
    // BWT callers whose output buffer is pinned host memory: the rows of the slots that round 0 settles are copied to
    // the host chunk by chunk while the remaining buckets are still being sorted (MSD path)
    u8 *h_U = nullptr;            // host address of the caller's U
    const u8 *h_T = nullptr;      // host address of the caller's T (its first symbols locate the row of suffix 0)
    u64 aux_r = 0;                // aux sampling: aux_I[p / r] = slot(p) + 1 for p % r == 0 (r a power of two)
    u32 *aux_I = nullptr;
};

struct SAResult {
    u32 *SA = nullptr;    // [n]   suffix array (nullptr when !want_sa)
    u32 *ISA = nullptr;   // [n]   inverse suffix array; complete only when isa_complete
    bool isa_complete = false;
    u64 primary = 0;      // slot of suffix 0, plus 1 (= the BWT primary index)
    // streamed rows (SAOptions::h_U): every slot outside [p0_lo, p0_hi) is already in the host buffer, shifted around the
    // dropped row; the slots of patch_slots[0..n_patch) were unresolved after round 0 and must be sent again
    bool u_streamed = false;
    u64 p0_lo = 0, p0_hi = 0;
    const u32 *patch_slots = nullptr; u64 n_patch = 0;
    void *scratch = nullptr;      // dead sort buffer, reusable by the caller after the build
    size_t scratch_bytes = 0;     // = 8n
};

// Device memory the SA core needs for a text of n symbols (excluding the text itself).
size_t sa_workspace_bytes(u64 n, int sym_bytes);

// Build SA and ISA of a device-resident text.  sym_bytes = 1 (bytes; the histogram is left in
// ctx.h_scalars[S_FREQ..+256) after the call), 4 (int32 symbols) or 8 (int64 symbols).
// Returns 0, or -2 on CUDA failure / exhausted workspace.  Arrays live in the ctx arena.
int build_sa(Ctx &c, const void *d_T, int sym_bytes, u64 n, const SAOptions &opt, SAResult *out);

// Post-processing stages (all device pointers).
//   bwt:   U[n] from the per-slot rows produced by build_sa (SAOptions::bwt_rows) and the primary index
int run_bwt_finish(Ctx &c, const u8 *d_T, const u8 *d_rows, u8 *d_U, u64 n, u64 primary);
//   streamed rows: rows of the slots listed in d_slots -> the caller's pinned U (device alias), shifted around the dropped row p0
int run_bwt_patch(Ctx &c, const u32 *d_slots, u64 count, const u8 *d_rows, u8 *U_host_dev, u64 p0);
//   plcp:  PLCP[n] from T (sym_bytes 1 or 4), SA.  d_T must be readable up to 16 bytes past the end.
size_t plcp_workspace_bytes(u64 n);
int run_plcp(Ctx &c, const void *d_T, int sym_bytes, const u32 *d_SA, u32 *d_PLCP, u64 n);
//   lcp:   LCP[i] = PLCP[SA[i]]
int run_lcp(Ctx &c, const u32 *d_PLCP, const u32 *d_SA, u32 *d_LCP, u64 n);
//   unbwt: text U[n] from BWT B[n] and the primary index
size_t unbwt_workspace_bytes(u64 n);
int run_unbwt(Ctx &c, const u8 *d_B, u8 *d_U, u64 n, u64 primary, u64 aux_r = 0, const u32 *d_I = nullptr, u64 n_aux = 0);   // d_I: aux samples (row of suffix j * aux_r)

// generalized suffix arrays: integer text with one distinct symbol per separator (gsa.cu)
size_t gsa_workspace_bytes(u64 n);
u32 *build_gsa_text(Ctx &c, const u8 *d_T, u64 n);

// building blocks of the distributed prefix doubling (sa_core.cu)
int dist_prepare(Ctx &c, const u8 *d_T, u64 n, int *k_out, int *K_out);
int dist_keys(Ctx &c, u64 lo, u64 count, u64 *d_keys, u32 *d_pos);
size_t rank_stage_workspace_bytes(u64 count);
int run_rank_stage(Ctx &c, const u64 *d_keys, const u32 *d_pos, const u32 *d_slot_in, u64 count, u32 slot_base,
                   u32 *d_sa_local, u32 *d_pair_pos, u32 *d_pair_rank, u32 *d_act_pos, u32 *d_act_slot, u32 *d_act_grp,
                   u64 *counts);
size_t sort_workspace_bytes(u64 count);
int run_sort_pairs(Ctx &c, u64 *ka, u32 *va, u64 *kb, u32 *vb, u64 count, int lo_bit, int hi_bit);
int run_sort_u32_pairs(Ctx &c, u32 *ka, u32 *va, u32 *kb, u32 *vb, u64 count, int lo_bit, int hi_bit);
void run_gather_u32(Ctx &c, const u32 *src, u64 src_len, const u32 *idx, u64 count, u32 idx_offset, u32 *out);
size_t route_workspace_bytes(u64 count);
int run_route(Ctx &c, const u32 *d_a, const u32 *d_b, u64 count, u64 add, u64 limit, u64 block, u32 world,
              u32 *d_a_out, u32 *d_b_out, u64 *counts_out);
int run_partition_by_splitters(Ctx &c, u64 *d_keys, u32 *d_pos, u64 count, const u64 *d_splitters, u32 nsplit,
                               u64 *d_keys_out, u32 *d_pos_out, u64 *counts_out);
size_t scatter_workspace_bytes(u64 count);
void run_scatter_u32(Ctx &c, u32 *dst, u64 dst_len, const u32 *idx, const u32 *val, u64 count, u32 idx_offset);

// exclusive max / sum / sum scans of the rank stage's tile aggregates tagg[3][ntiles]; totals of the sums -> out_counts[0..1]
void run_rank_scan(Ctx &c, u32 *tagg, u64 ntiles, u64 *out_counts);

// distributed prefix doubling over several GPUs of one node, 64-bit positions (dist64.cu)
int sa64_multi(const u8 *T, i64 *SA, u64 n, i64 *freq, const int *devices, int ndev, void *stats /* libsais_cuda_dist_stats, nullable */);
i64 bwt64_multi(const u8 *T, u8 *U, i64 *A, u64 n, i64 *freq, u64 aux_r, i64 *aux_I, const int *devices, int ndev);

// 16-bit symbols (post.cu, gsa.cu): widened text for the SA / PLCP cores, 65536-bin frequencies, BWT rows from the SA,
// inverse BWT over a 65536-symbol alphabet
void run_widen16(Ctx &c, const uint16_t *src, u32 *dst, u64 n);
void run_hist_u16(Ctx &c, const uint16_t *d_T, u64 n, u64 *d_hist /* 65536 */);
int run_bwt16(Ctx &c, const uint16_t *d_T, const u32 *d_SA, uint16_t *d_U, u64 n, u64 aux_r, u32 *d_I, u64 *primary_out);
size_t unbwt16_workspace_bytes(u64 n);
int run_unbwt16(Ctx &c, uint16_t *d_B /* clobbered */, uint16_t *d_U, u64 n, u64 primary, u64 aux_r, const u32 *d_I, u64 n_aux);
u32 *build_gsa_text16(Ctx &c, const uint16_t *d_T, u64 n);

// conversions used by the 64-bit API
void run_widen(Ctx &c, const u32 *src, i64 *dst, u64 n);
void run_narrow(Ctx &c, const i64 *src, u32 *dst, u64 n);

void run_byte_histogram(Ctx &c, const u8 *d_T, u64 n);   // -> d_scalars[S_FREQ..+256)

}  // namespace lsc
