"""Workload for the per-kernel ncu captures (profiles/kernels_r2.md): one pass over every kernel class of the library.
  c2  : libsais_bwt on 256 MiB random bytes (MSD round 0: hist16, both partition levels, bucket_sort, rank stage, bwt_finish)
  c3s : libsais + plcp + lcp + bwt + unbwt on config 3 at 1/10 scale (LSD round 0, local_count / local_sort rounds,
        partitioned scatter, phi, PLCP levels + chunk kernel, LCP gather, unBWT walks and list ranking)
usage: python tools/prof_kernels.py [c2] [c3s]"""
import sys

import torch

sys.path.insert(0, ".")
import libsais_b200
from libsais_b200 import gen

which = sys.argv[1:] or ["c2", "c3s"]
ctx = libsais_b200.Context(0)
if "c2" in which:
    n = 1 << 28
    dT = gen.rand_bytes_torch(2, n)
    dU = torch.empty(n, dtype=torch.uint8, device="cuda")
    for _ in range(1):
        assert ctx.bwt_dev(dT.data_ptr(), dU.data_ptr(), n) > 0
    del dT, dU
if "c3s" in which:
    dT = gen.repetitive_dna_torch(1_900_000, 100)
    n = dT.numel()
    dSA = torch.empty(n, dtype=torch.int32, device="cuda"); dP = torch.empty(n, dtype=torch.int32, device="cuda")
    dL = torch.empty(n, dtype=torch.int32, device="cuda"); dU = torch.empty(n, dtype=torch.uint8, device="cuda"); dB = torch.empty(n, dtype=torch.uint8, device="cuda")
    assert ctx.sa_dev(dT.data_ptr(), dSA.data_ptr(), n) == 0
    assert ctx.plcp_dev(dT.data_ptr(), dSA.data_ptr(), dP.data_ptr(), n) == 0
    assert ctx.lcp_dev(dP.data_ptr(), dSA.data_ptr(), dL.data_ptr(), n) == 0
    pr = ctx.bwt_dev(dT.data_ptr(), dU.data_ptr(), n)
    assert pr > 0 and ctx.unbwt_dev(dU.data_ptr(), dB.data_ptr(), n, pr) == 0
if "c3d" in which:                                  # config 3 at full size, SA only (the doubling rounds)
    dT = gen.repetitive_dna_torch(19_000_000, 100)
    n = dT.numel()
    dSA = torch.empty(n, dtype=torch.int32, device="cuda")
    assert ctx.sa_dev(dT.data_ptr(), dSA.data_ptr(), n) == 0
torch.cuda.synchronize()
print("prof_kernels done")
