"""Differential check of the round-0 MSD path (partition.cuh) against the compiled reference / oracle:
forces LIBSAIS_CUDA_MSD=2 (MSD path for every text it can handle) and compares SA and BWT with the
stable LSD path (LIBSAIS_CUDA_MSD=0) and the CPU result.  usage: python tools/msd_check.py [--small]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import _libs
from libsais_b200 import gen


def texts(small):
    rng = np.random.default_rng(11)
    out = [("rand100", gen.rand_bytes(3, 100)), ("rand5000", gen.rand_bytes(4, 5000)), ("rand70k", gen.rand_bytes(5, 70_000)),
           ("dna4609", gen.dna(6, 4609)), ("dna300k", gen.dna(7, 300_000)),
           ("sigma16", (rng.integers(0, 16, 200_000) + 97).astype(np.uint8)),
           ("sigma2", (rng.integers(0, 2, 100_000) + 48).astype(np.uint8)),
           ("zero_tail", np.concatenate([gen.rand_bytes(8, 50_000), np.zeros(37, dtype=np.uint8)])),
           ("repeat", np.tile(gen.rand_bytes(9, 20_000), 5))]
    if not small:
        out += [("rand1M", gen.rand_bytes(2, 1 << 20)), ("rand16M", gen.rand_bytes(2, 1 << 24)), ("dna8M", gen.dna(1, 1 << 23)),
                ("rand64M", gen.rand_bytes(2, 1 << 26))]
    return out


def main():
    small = "--small" in sys.argv
    cu, ref = _libs.cuda(), (_libs.ref() or _libs.oracle())
    bad = 0
    for name, T in texts(small):
        t0 = time.time()
        rs, SAr = ref.sa(T)
        rb, Ur = ref.bwt(T)
        tr = time.time() - t0
        for mode in ("2", "0"):
            os.environ["LIBSAIS_CUDA_MSD"] = mode
            rc, SA = cu.sa(T)
            rcb, U = cu.bwt(T)
            ok_sa = rc == 0 and bool((SA == SAr).all())
            ok_bwt = rcb == rb and bool((U == Ur).all())
            msg = "ok" if ok_sa and ok_bwt else "MISMATCH"
            if not ok_sa and rc == 0:
                i = int(np.argmax(SA != SAr))
                msg += " SA first diff at %d: got %d want %d (n=%d)" % (i, SA[i], SAr[i], len(T))
            if not ok_bwt:
                msg += " BWT rc %d vs %d" % (rcb, rb)
            if rc != 0:
                msg += " sa rc %d" % rc
            print("%-10s n=%-9d msd=%s %s (cpu %.2fs)" % (name, len(T), mode, msg, tr), flush=True)
            bad += 0 if (ok_sa and ok_bwt) else 1
    os.environ.pop("LIBSAIS_CUDA_MSD", None)
    print("msd_check: %d failures" % bad)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
