"""Distributed suffix array driver (torchrun): builds the SA of one text across all ranks with
libsais_b200/dist.py, verifies each rank's slice against the single-GPU SA (when it fits) and
prints timing.  usage: torchrun --nproc-per-node G tools/dist_sa.py <log2n|n> [kind] [--no-verify]"""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, ".")
import libsais_b200
from libsais_b200 import gen
from libsais_b200.dist import DistributedSA, verify_distributed


def main():
    arg = sys.argv[1] if len(sys.argv) > 1 else "24"
    n = (1 << int(arg)) if int(arg) < 64 else int(arg)
    kind = sys.argv[2] if len(sys.argv) > 2 else "dna"
    verify = "--no-verify" not in sys.argv
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29533")
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
    dT = None
    if kind == "dna" and n >= (1 << 27):
        dT = gen.dna_torch(5, n, device="cuda")        # same text as gen.dna(5, n), generated on the device
    elif kind == "dna":
        T = gen.dna(5, n)
    elif kind == "bytes":
        T = gen.rand_bytes(2, n)
    elif kind == "rep":
        T = gen.repetitive_dna(n // 50, 50); n = len(T)
    elif kind == "zeros":
        T = np.zeros(n, dtype=np.uint8)
    else:
        T = np.resize(np.frombuffer(b"abracadabra", dtype=np.uint8), n)
    if dT is None:
        dT = torch.from_numpy(T).cuda()
    ctx = libsais_b200.Context(local)
    if "--no-warmup" not in sys.argv:
        d = DistributedSA(ctx, dT, n)
        d.run()                                  # warm-up (workspace growth, NCCL channels)
        del d
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    d = DistributedSA(ctx, dT, n)
    sa, base = d.run()
    torch.cuda.synchronize(); dist.barrier()
    dt = time.perf_counter() - t0
    ok = None
    vinfo = None
    if "--verify-dist" in sys.argv:
        okd, vinfo = verify_distributed(d)
        vinfo["ok"] = okd
    if verify and "--verify-dist" not in sys.argv:
        full = torch.empty(n, dtype=torch.int32, device="cuda")
        assert ctx.sa_dev(dT.data_ptr(), full.data_ptr(), n) == 0
        ok = bool(torch.equal(full[base: base + sa.numel()], sa))
        if not ok:
            ref = full[base: base + sa.numel()]
            bad = (ref != sa).nonzero().flatten()
            print("rank", rank, "base", base, "m", sa.numel(), "mismatches", bad.numel(), "first", bad[:5].tolist(),
                  "got", sa[bad[:5]].tolist(), "want", ref[bad[:5]].tolist(), flush=True)
        flag = torch.tensor([1 if ok else 0], device="cuda"); dist.all_reduce(flag, op=dist.ReduceOp.MIN); ok = bool(flag.item())
    sizes = [None] * world
    dist.all_gather_object(sizes, int(sa.numel()))
    if rank == 0:
        print(json.dumps({"n": n, "kind": kind, "world": world, "seconds": round(dt, 4), "mbs": round(n / 1e6 / dt, 1), "parity_vs_single_gpu": ok, "distributed_check": vinfo,
                          "gpu_mem_gb": round(torch.cuda.max_memory_allocated() / 2**30, 1), "phases_ms_rank0": getattr(d, "phases", None) or None,
                          "slice_sizes": sizes, "rounds": [(r["h"], r["local"], r["active_local"]) for r in d.rounds]}), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
