"""Distributed prefix doubling: suffix array of ONE text across several GPUs (BASELINE config 5's
mechanism).  One process per GPU; torch.distributed (NCCL over
NVLink) carries the exchanges; all sorting / ranking is done by the library's own kernels through
the device-pointer building blocks of include/libsais_cuda.h.

Layout: the text is replicated on every rank.  Rank r owns
  * positions [r*B, (r+1)*B)  -> its slice of ISA (rank of every suffix that starts there), and
  * after the round-0 sample sort, a contiguous range of KEYS -> a contiguous slice of the sorted
    order (global slots [base_r, base_r + m_r)): its slice of the final suffix array.
Equal keys never straddle ranks, so every group of tied suffixes lives on one rank and the
single-GPU rank stage applies unchanged.  Per round there are three all-to-alls:
  requests (p + h) to the position owners, their ISA values back, and the new (position, rank)
  pairs to the position owners.
Round-1 limit: positions are u32 (n < 2^32 - 16): texts larger than one GPU's ~50 B/suffix working
set but below 4 Gi symbols.  2^34 symbols (config 5 proper) needs 64-bit positions -- round 2.
"""
import ctypes as C

import torch
import torch.distributed as dist

from . import load_library

_M32 = 0xFFFFFFFF


def _u32(t):
    """int32 tensor holding u32 bit patterns -> non-negative int64."""
    return t.long() & _M32


def _bits(x):
    b = 0
    while x:
        b += 1
        x >>= 1
    return max(b, 1)


class _Lib:
    def __init__(self):
        lib = load_library()
        vp, i32, i64, u32 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint32
        lib.libsais_cuda_dist_prepare.restype = i64
        lib.libsais_cuda_dist_prepare.argtypes = [vp, vp, i64, C.POINTER(i32), C.POINTER(i32)]
        lib.libsais_cuda_dist_keys.restype = i64
        lib.libsais_cuda_dist_keys.argtypes = [vp, i64, i64, vp, vp]
        lib.libsais_cuda_sort_pairs_dev.restype = i64
        lib.libsais_cuda_sort_pairs_dev.argtypes = [vp, vp, vp, vp, vp, i64, i32, i32]
        lib.libsais_cuda_sort_u32_pairs_dev.restype = i64
        lib.libsais_cuda_sort_u32_pairs_dev.argtypes = [vp, vp, vp, vp, vp, i64, i32, i32]
        lib.libsais_cuda_rank_stage_dev.restype = i64
        lib.libsais_cuda_rank_stage_dev.argtypes = [vp, vp, vp, vp, i64, u32, vp, vp, vp, vp, vp, vp, C.POINTER(C.c_uint64)]
        lib.libsais_cuda_gather_u32_dev.restype = i64
        lib.libsais_cuda_gather_u32_dev.argtypes = [vp, vp, i64, vp, i64, u32, vp]
        lib.libsais_cuda_scatter_u32_dev.restype = i64
        lib.libsais_cuda_scatter_u32_dev.argtypes = [vp, vp, i64, vp, vp, i64, u32]
        lib.libsais_cuda_dist_route_dev.restype = i64
        lib.libsais_cuda_dist_route_dev.argtypes = [vp, vp, vp, i64, i64, i64, i64, i32, vp, vp, C.POINTER(C.c_uint64)]
        lib.libsais_cuda_dist_partition_dev.restype = i64
        lib.libsais_cuda_dist_partition_dev.argtypes = [vp, vp, vp, i64, vp, i32, vp, vp, C.POINTER(C.c_uint64)]
        self.lib = lib


def _check(rc, what):
    if rc < 0:
        raise RuntimeError("libsais_cuda %s failed: %d" % (what, rc))
    return rc


def _exchange(send, send_counts, world):
    """all_to_all of variable-sized segments of `send` (segments ordered by destination)."""
    sc = torch.tensor(send_counts, dtype=torch.int64, device=send.device)
    rc = torch.empty_like(sc)
    dist.all_to_all_single(rc, sc)
    recv_counts = [int(x) for x in rc.tolist()]
    out = torch.empty(sum(recv_counts), dtype=send.dtype, device=send.device)
    dist.all_to_all_single(out, send, output_split_sizes=recv_counts, input_split_sizes=list(send_counts))
    return out, recv_counts


def _exchange_known(send, send_counts, recv_counts):
    out = torch.empty(sum(recv_counts), dtype=send.dtype, device=send.device)
    dist.all_to_all_single(out, send, output_split_sizes=list(recv_counts), input_split_sizes=list(send_counts))
    return out


class DistributedSA:
    """Builds the suffix array of a replicated device text `dT` (uint8, n symbols) across the
    process group.  After run(): self.sa_local (int32 tensor, u32 bit patterns) holds global
    slots [self.base, self.base + len(self.sa_local)) of the suffix array."""

    def __init__(self, ctx, dT, n, samples=8192):
        self.L = _Lib().lib
        self.ctx, self.h = ctx, ctx.handle
        self.dT, self.n = dT, int(n)
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.dev = dT.device
        self.samples = samples
        self.B = (self.n + self.world - 1) // self.world          # positions per owner
        self.lo = min(self.n, self.rank * self.B)
        self.hi = min(self.n, self.lo + self.B)
        self.rounds = []

    # ---- helpers -------------------------------------------------------------------------
    def _sort_pairs(self, keys, vals, lo_bit, hi_bit):
        ka, va = torch.empty_like(keys), torch.empty_like(vals)
        w = _check(self.L.libsais_cuda_sort_pairs_dev(self.h, keys.data_ptr(), vals.data_ptr(), ka.data_ptr(), va.data_ptr(),
                                                      keys.numel(), lo_bit, hi_bit), "sort_pairs")
        return (ka, va) if w == 1 else (keys, vals)

    def _route(self, a, b, add=0):
        """Group (a, b) items (int32 tensors, u32 bit patterns) by the owner of position a + add with the library's
        routing kernels (one onesweep digit pass); items with a + add >= n are dropped.  Returns a + add and b in
        destination order (only the routed items) and the per-destination counts."""
        nitems = a.numel()
        a_out = torch.empty(max(nitems, 1), dtype=torch.int32, device=self.dev)
        b_out = torch.empty(max(nitems, 1), dtype=torch.int32, device=self.dev)
        counts = (C.c_uint64 * 64)()
        _check(self.L.libsais_cuda_dist_route_dev(self.h, a.data_ptr(), b.data_ptr(), nitems, add, self.n, self.B, self.world,
                                                  a_out.data_ptr(), b_out.data_ptr(), counts), "route")
        counts = [int(counts[i]) for i in range(self.world)]
        tot = sum(counts)
        return a_out[:tot], b_out[:tot], counts

    def _scatter_isa(self, pos, rank):
        """Route (position, rank) pairs to the owners of the positions and store them in the ISA slices."""
        pos_s, rank_s, counts = self._route(pos, rank)
        pos_r, rc = _exchange(pos_s, counts, self.world)
        rank_r = _exchange_known(rank_s, counts, rc)
        if pos_r.numel():
            _check(self.L.libsais_cuda_scatter_u32_dev(self.h, self.isa.data_ptr(), self.isa.numel(), pos_r.data_ptr(), rank_r.data_ptr(),
                                                       pos_r.numel(), self.lo), "scatter")

    # ---- the algorithm -------------------------------------------------------------------
    def run(self):
        # torch's glue ops and the NCCL collectives are issued on the context's own stream, so they are
        # ordered with the library's kernels without host synchronisation
        torch.cuda.current_stream(self.dev).synchronize()
        with torch.cuda.stream(torch.cuda.ExternalStream(self.ctx.stream, device=self.dev)):
            out = self._run()
        torch.cuda.current_stream(self.dev).wait_stream(torch.cuda.ExternalStream(self.ctx.stream, device=self.dev))
        return out

    def _tick(self, name):
        if not self.timing:
            return
        import time
        torch.cuda.synchronize()
        t = time.perf_counter()
        self.phases.append((name, round((t - self._t) * 1e3, 2)))
        self._t = t

    def _run(self):
        import os, time
        self.timing = bool(os.environ.get("LSC_DIST_TIMING"))
        self.phases = []
        if self.timing:
            torch.cuda.synchronize()
        self._t = time.perf_counter()
        n, world, dev = self.n, self.world, self.dev
        k_out, kb_out = C.c_int32(0), C.c_int32(0)
        _check(self.L.libsais_cuda_dist_prepare(self.h, self.dT.data_ptr(), n, C.byref(k_out), C.byref(kb_out)), "dist_prepare")
        k, K = k_out.value, kb_out.value
        self._tick("prepare_call")
        key_bits = K + _bits(k)                 # k-mer + length field (see dist_keys_kernel)
        cnt = self.hi - self.lo
        self.isa = torch.zeros(max(cnt, 1), dtype=torch.int32, device=dev)

        # ---- round 0: keys of the owned positions, sample sort over the ranks
        keys = torch.empty(max(cnt, 1), dtype=torch.int64, device=dev)[:cnt]
        pos = torch.empty(max(cnt, 1), dtype=torch.int32, device=dev)[:cnt]
        self._tick("alloc_round0")
        _check(self.L.libsais_cuda_dist_keys(self.h, self.lo, cnt, keys.data_ptr(), pos.data_ptr()), "dist_keys")
        self._tick("keys")
        # splitters from a random sample of the (unsorted) keys; keys are < 2^55: signed order = unsigned order
        g = torch.Generator(device=dev); g.manual_seed(1234 + self.rank)
        if cnt:
            mine = keys[torch.randint(0, cnt, (self.samples,), device=dev, generator=g)]
        else:
            mine = torch.full((self.samples,), (1 << 62), dtype=torch.int64, device=dev)
        allsamp = torch.empty(self.samples * world, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(allsamp, mine.contiguous())
        allsamp, _ = torch.sort(allsamp)                                  # 8192*G values: plumbing, not the data path
        splitters = allsamp[[(i * allsamp.numel()) // world for i in range(1, world)]] if world > 1 else allsamp[:0]
        # destination of a key = number of splitters <= key (equal keys share a destination); it rides in the
        # key's free top byte, so ONE onesweep digit pass groups the (key, position) pairs by destination
        keys_p, pos_p = torch.empty_like(keys), torch.empty_like(pos)
        cbuf = (C.c_uint64 * 65)()
        _check(self.L.libsais_cuda_dist_partition_dev(self.h, keys.data_ptr(), pos.data_ptr(), cnt, splitters.contiguous().data_ptr(), world - 1,
                                                      keys_p.data_ptr(), pos_p.data_ptr(), cbuf), "partition")
        keys, pos = keys_p, pos_p
        counts = [int(cbuf[i]) for i in range(world)]
        self._tick("splitters")
        keys_r, rc = _exchange(keys, counts, world)
        pos_r = _exchange_known(pos, counts, rc)
        self._tick("all_to_all_keys")
        m = keys_r.numel()
        keys_r, pos_r = self._sort_pairs(keys_r, pos_r, 0, key_bits)      # merge of the received sorted runs
        self._tick("merge_sort")
        sizes = torch.empty(world, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(sizes, torch.tensor([m], dtype=torch.int64, device=dev))
        sizes = sizes.tolist()
        self.base = int(sum(sizes[: self.rank]))
        self.sa_local = torch.empty(max(m, 1), dtype=torch.int32, device=dev)[:m]

        act_pos, act_slot, act_grp, n_act, n_grp = self._rank_stage(keys_r, pos_r, None, m)
        self._tick("rank_stage0+isa")
        del keys_r, pos_r, keys, pos
        self.rounds.append({"h": 0, "local": m, "active_local": n_act})

        # ---- doubling rounds
        rank_bits = _bits(n)
        h = k
        while True:
            tot = torch.tensor([n_act], dtype=torch.int64, device=dev)
            dist.all_reduce(tot)
            if int(tot.item()) == 0:
                break
            # k2 = ISA[p + h] + 1 (0 past the end): request / response all-to-all with the position owners
            ident = torch.arange(n_act, dtype=torch.int32, device=dev)
            q_s, id_s, counts = self._route(act_pos, ident, add=h)
            nreq = sum(counts)
            req, rc = _exchange(q_s, counts, world)
            ans = torch.empty(max(req.numel(), 1), dtype=torch.int32, device=dev)[: req.numel()]
            if req.numel():
                _check(self.L.libsais_cuda_gather_u32_dev(self.h, self.isa.data_ptr(), self.isa.numel(), req.data_ptr(), req.numel(),
                                                          self.lo, ans.data_ptr()), "gather")
            resp = _exchange_known(ans, rc, counts)
            k2 = torch.zeros(max(n_act, 1), dtype=torch.int64, device=dev)[:n_act]
            if nreq:
                k2[id_s.long()] = _u32(resp) + 1
            grp_bits = _bits(max(n_grp - 1, 1))
            keys = (_u32(act_grp) << rank_bits) | k2
            keys, spos = self._sort_pairs(keys, act_pos.clone(), 0, rank_bits + grp_bits)
            act_pos, act_slot, act_grp, n_act_new, n_grp = self._rank_stage(keys, spos, act_slot, n_act)
            self.rounds.append({"h": h, "local": n_act, "active_local": n_act_new})
            n_act = n_act_new
            self._tick("round_h%d" % h)
            h *= 2
            if len(self.rounds) > 80:
                raise RuntimeError("distributed prefix doubling did not converge")
        return self.sa_local, self.base

    def _rank_stage(self, keys, pos, slot_in, count):
        dev = self.dev
        pair_pos = torch.empty(max(count, 1), dtype=torch.int32, device=dev)
        pair_rank = torch.empty(max(count, 1), dtype=torch.int32, device=dev)
        a_pos = torch.empty(max(count, 1), dtype=torch.int32, device=dev)
        a_slot = torch.empty(max(count, 1), dtype=torch.int32, device=dev)
        a_grp = torch.empty(max(count, 1), dtype=torch.int32, device=dev)
        counts = (C.c_uint64 * 2)()
        _check(self.L.libsais_cuda_rank_stage_dev(self.h, keys.data_ptr(), pos.data_ptr(), slot_in.data_ptr() if slot_in is not None else None,
                                                  count, self.base & _M32, self.sa_local.data_ptr(), pair_pos.data_ptr(), pair_rank.data_ptr(),
                                                  a_pos.data_ptr(), a_slot.data_ptr(), a_grp.data_ptr(), counts), "rank_stage")
        n_act, n_grp = int(counts[0]), int(counts[1])
        self._scatter_isa(pair_pos[:count], pair_rank[:count])
        return a_pos[:n_act].clone(), a_slot[:n_act].clone(), a_grp[:n_act].clone(), n_act, n_grp


def verify_distributed(d, pairs=200000, depth=256, seed=1):
    """Checks a DistributedSA result without a single-GPU reference: (1) the slices together are a
    permutation of [0, n) (positions routed to their owners, every owner sees each of its positions
    exactly once); (2) sampled adjacent suffixes inside every slice are in order (direct comparison of
    up to `depth` symbols of the replicated text; undecided pairs are reported, not failed).
    Returns (ok, info) on every rank."""
    dev, n, world = d.dev, d.n, d.world
    with torch.cuda.stream(torch.cuda.ExternalStream(d.ctx.stream, device=dev)):
        sa = d.sa_local
        pos_s, _unused, counts = d._route(sa, sa)
        got, _ = _exchange(pos_s, counts, world)
        cnt = d.hi - d.lo
        seen = torch.zeros(max(cnt, 1), dtype=torch.int32, device=dev)
        if got.numel():
            seen.index_put_(((_u32(got) - d.lo),), torch.ones(got.numel(), dtype=torch.int32, device=dev), accumulate=True)
        perm_ok = bool(got.numel() == cnt and (cnt == 0 or bool((seen[:cnt] == 1).all())))
        m = sa.numel()
        bad = undecided = 0
        if m > 1:
            g = torch.Generator(device=dev); g.manual_seed(seed + d.rank)
            i = torch.randint(1, m, (min(pairs, m - 1),), device=dev, generator=g)
            a, b = _u32(sa[i - 1]), _u32(sa[i])
            off = torch.arange(depth, device=dev, dtype=torch.int64)
            ia, ib = a[:, None] + off[None, :], b[:, None] + off[None, :]
            ta = torch.where(ia < n, d.dT[torch.clamp(ia, max=n - 1)].to(torch.int16), torch.full_like(ia, -1, dtype=torch.int16))
            tb = torch.where(ib < n, d.dT[torch.clamp(ib, max=n - 1)].to(torch.int16), torch.full_like(ib, -1, dtype=torch.int16))
            diff = ta != tb
            anyd = diff.any(dim=1)
            first = torch.argmax(diff.to(torch.int8), dim=1)
            va = ta.gather(1, first[:, None]).squeeze(1); vb = tb.gather(1, first[:, None]).squeeze(1)
            bad = int((anyd & (va > vb)).sum())
            undecided = int((~anyd).sum())
        flags = torch.tensor([1 if perm_ok else 0, bad, undecided], dtype=torch.int64, device=dev)
        mins = flags.clone(); dist.all_reduce(mins, op=dist.ReduceOp.MIN)
        sums = flags.clone(); dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    torch.cuda.synchronize()
    ok = bool(mins[0].item() == 1 and sums[1].item() == 0)
    return ok, {"permutation": bool(mins[0].item() == 1), "order_violations": int(sums[1].item()), "undecided_pairs": int(sums[2].item()),
                "pairs_per_rank": pairs, "depth": depth}


def distributed_suffix_array(ctx, dT, n):
    """Convenience wrapper: returns (sa_local, base, rounds)."""
    d = DistributedSA(ctx, dT, n)
    sa, base = d.run()
    return sa, base, d.rounds
