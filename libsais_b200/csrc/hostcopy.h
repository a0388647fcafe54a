// hostcopy.h -- the host side of the boundary: moving the caller's buffers to
// and from the GPU.  Callers of the reference API pass ordinary malloc'd memory; a plain
// cudaMemcpyAsync from pageable memory is staged by the driver at ~12 GB/s on this box (a 256 MiB
// BWT then spends 46 ms in copies against 15 ms of kernels).  Large pageable transfers therefore go
// through a small ring of pinned chunks owned by the context: a few host threads memcpy chunks
// between the caller's buffer and the ring while the DMA engine moves the previous chunks.
// Pinned (page-locked / registered) caller buffers are copied directly.
#pragma once
#include "ctx.h"

namespace lsc {

// Both calls are ordered with respect to the context's stream: h2d returns after the data is
// queued ahead of any later work on c.stream; d2h waits for the work already on c.stream first.
// They return false on a CUDA failure (recorded in the context).
bool copy_h2d(Ctx &c, void *d_dst, const void *h_src, size_t bytes);
bool copy_d2h(Ctx &c, void *h_dst, const void *d_src, size_t bytes);
// page-locked (cudaHostAlloc / cudaHostRegister) host memory?
bool host_is_pinned(const void *p);

}  // namespace lsc
