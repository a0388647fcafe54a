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
    S_TICKET  = 264,   // [264, 272): tickets of the rank kernels (u32 views)
    S_MISC    = 272
};

struct SAResult {
    u32 *SA = nullptr;    // [n]   suffix array
    u32 *ISA = nullptr;   // [n]   inverse suffix array (rank of every suffix)
    void *scratch = nullptr;      // dead sort buffer, reusable by the caller after the build
    size_t scratch_bytes = 0;     // = 8n
};

// Device memory the SA core needs for a text of n symbols (excluding the text itself).
size_t sa_workspace_bytes(u64 n, int sym_bytes);

// Build SA and ISA of a device-resident text.  sym_bytes = 1 (bytes; the histogram is left in
// ctx.h_scalars[S_FREQ..+256) after the call), 4 (int32 symbols) or 8 (int64 symbols).
// Returns 0, or -2 on CUDA failure / exhausted workspace.  Arrays live in the ctx arena.
// sa_out (optional): caller's device buffer for SA instead of an arena array.
int build_sa(Ctx &c, const void *d_T, int sym_bytes, u64 n, u32 *sa_out, SAResult *out);

// Post-processing stages (all device pointers).
//   bwt:   U[n] from T, SA, ISA;  *primary (host) = ISA[0]+1;  aux (device, optional): I[j] = ISA[j*r]+1
int run_bwt(Ctx &c, const u8 *d_T, const u32 *d_SA, const u32 *d_ISA, u8 *d_U, u64 n,
            u64 r, u32 *d_I, u64 n_aux);
//   plcp:  PLCP[n] from T (sym_bytes 1 or 4), SA.  d_T must be readable up to 16 bytes past the end.
int run_plcp(Ctx &c, const void *d_T, int sym_bytes, const u32 *d_SA, u32 *d_PLCP, u64 n);
//   lcp:   LCP[i] = PLCP[SA[i]]
int run_lcp(Ctx &c, const u32 *d_PLCP, const u32 *d_SA, u32 *d_LCP, u64 n);
//   unbwt: text U[n] from BWT B[n] and the primary index
size_t unbwt_workspace_bytes(u64 n);
int run_unbwt(Ctx &c, const u8 *d_B, u8 *d_U, u64 n, u64 primary);

// conversions used by the 64-bit API
void run_widen(Ctx &c, const u32 *src, i64 *dst, u64 n);
void run_narrow(Ctx &c, const i64 *src, u32 *dst, u64 n);

void run_byte_histogram(Ctx &c, const u8 *d_T, u64 n);   // -> d_scalars[S_FREQ..+256)

}  // namespace lsc
