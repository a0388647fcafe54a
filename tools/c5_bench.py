"""BASELINE config 5: libsais64-class suffix array of an iid ACGT text (seed 5) far beyond one GPU's working set, by the
in-library distributed prefix doubling (dist64.cu) -- weak scaling 2^31 symbols per GPU: 2 GPUs 2^32, 4 GPUs 2^33,
8 GPUs 2^34 (16 GiB text).  The result stays distributed on the GPUs (the host copy of a 128 GiB SA is not the
measured path) and is PROVEN by the distributed checker (ISA[SA[i]] == i and the neighbour-order rule for every slot).
usage: python tools/c5_bench.py [log2 symbols per GPU, default 31] [G ...]"""
import ctypes as C
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, "."); sys.path.insert(0, "tools")
import libsais_b200
from libsais_b200 import gen
from dist64_check import DistStats


def main():
    per = int(sys.argv[1]) if len(sys.argv) > 1 else 31
    ndev = torch.cuda.device_count()
    Gs = [int(a) for a in sys.argv[2:]] or [g for g in (2, 4, 8) if g <= ndev]
    lib = libsais_b200.load_library()
    lib.libsais_cuda_sa64_multi.restype = C.c_int64
    nmax = (1 << per) * max(Gs)
    t0 = time.time()
    T = np.empty(nmax, dtype=np.uint8)
    ch = 1 << 28
    for lo in range(0, nmax, ch):                      # same text as gen.dna(5, n): generated on GPU 0, chunk by chunk
        hi = min(nmax, lo + ch)
        T[lo:hi] = gen.dna_torch_range(5, lo, hi, device="cuda:0").cpu().numpy()
    torch.cuda.empty_cache()
    print(json.dumps({"generated_symbols": nmax, "gen_s": round(time.time() - t0, 1)}), flush=True)

    def run(n, G, verify):
        st = DistStats(); st.verify = 1 if verify else 0
        t0 = time.time()
        rc = lib.libsais_cuda_sa64_multi(T.ctypes.data_as(C.c_void_p), None, C.c_int64(n), None, None, C.c_int32(G), C.byref(st))
        return rc, st, time.time() - t0

    run(1 << 24, max(Gs), False)                       # warm-up: peer mappings, module loading
    for G in Gs:
        n = (1 << per) * G
        rc, st, wall = run(n, G, True)
        rec = {"config": "c5", "n": n, "log2n": per + int(np.log2(G)), "gpus": G, "rc": int(rc), "device_s": round(st.seconds_device, 3),
               "wall_s_incl_upload_and_check": round(wall, 2), "mbs_device": round(n / 1e6 / max(st.seconds_device, 1e-9), 1),
               "rounds": st.rounds, "key_symbols": st.key_symbols, "slice_max": st.slice_max, "unresolved_after_round0": st.active_after_round0,
               "nvlink_bytes": st.exchanged_bytes, "nvlink_gbs_per_gpu_avg_over_run": round(st.exchanged_bytes / G / max(st.seconds_device, 1e-9) / 1e9, 1),
               "proven_by_distributed_checker": st.verify == 1, "violations": st.verify_violations,
               "phase_s": {k: round(st.phase_seconds[i], 3) for i, k in enumerate(("keys_route", "local_sort", "rank_stage", "isa_scatter", "later_rounds"))}}
        print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
