"""Parity / timing of the in-library distributed prefix doubling (dist64.cu) through libsais_cuda_sa64_multi:
G ranks (one host thread each) on the visible GPUs -- ranks share GPUs when there are fewer -- against the CPU
reference.  usage: python tools/dist64_check.py [G ...] [--big log2n]"""
import ctypes as C
import json
import sys
import time

import numpy as np

sys.path.insert(0, "."); sys.path.insert(0, "tests")
import _libs
import libsais_b200
from libsais_b200 import gen


class DistStats(C.Structure):
    _fields_ = [("n_gpus", C.c_int32), ("rounds", C.c_int32), ("key_symbols", C.c_int32), ("key_bits", C.c_int32),
                ("slice_max", C.c_uint64), ("active_after_round0", C.c_uint64), ("exchanged_bytes", C.c_uint64),
                ("seconds_total", C.c_double), ("seconds_device", C.c_double), ("verify", C.c_int32), ("reserved", C.c_int32),
                ("verify_violations", C.c_uint64), ("phase_seconds", C.c_double * 8)]


def run(lib, T, G, want_sa=True, verify=True):
    n = len(T)
    SA = np.full(n, -1, dtype=np.int64) if want_sa else None
    freq = np.zeros(256, dtype=np.int64)
    st = DistStats()
    st.verify = 1 if verify else 0
    lib.libsais_cuda_sa64_multi.restype = C.c_int64
    rc = lib.libsais_cuda_sa64_multi(T.ctypes.data_as(C.c_void_p), None if SA is None else SA.ctypes.data_as(C.c_void_p), C.c_int64(n),
                                     freq.ctypes.data_as(C.c_void_p), None, C.c_int32(G), C.byref(st))
    return rc, SA, freq, st


def main():
    lib = libsais_b200.load_library()
    ref = _libs.ref() or _libs.oracle()
    args = sys.argv[1:]
    if "--big" in args:
        args = args[:args.index("--big")] + args[args.index("--big") + 2:]
    Gs = [int(a) for a in args if a.isdigit()] or [1, 2, 3]
    texts = {"dna1M": gen.dna(5, 1 << 20), "bytes300k": gen.rand_bytes(2, 300_007), "rep1.5M": gen.repetitive_dna(30_000, 50),
             "zeros50k": np.zeros(50_000, dtype=np.uint8), "abra": np.resize(np.frombuffer(b"abracadabra", dtype=np.uint8), 300_007),
             "tiny": np.frombuffer(b"mississippi", dtype=np.uint8).copy()}
    bad = 0
    for name, T in texts.items():
        want = ref.sa(T, 64)[1]
        for G in Gs:
            rc, SA, freq, st = run(lib, T, G)
            ok = rc == 0 and bool((SA == want).all()) and bool((freq == np.bincount(T, minlength=256)).all()) and st.verify == 1
            bad += 0 if ok else 1
            msg = ""
            if rc == 0 and not ok:
                i = int(np.argmax(SA != want)); msg = " first diff at %d: got %d want %d" % (i, SA[i], want[i])
            print(json.dumps({"text": name, "n": len(T), "G": G, "rc": int(rc), "ok": ok, "rounds": st.rounds, "k": st.key_symbols,
                              "active0": st.active_after_round0, "exchanged_MB": round(st.exchanged_bytes / 1e6, 1),
                              "device_s": round(st.seconds_device, 4)}) + msg, flush=True)
    if "--big" in sys.argv:
        lg = int(sys.argv[sys.argv.index("--big") + 1])
        import torch
        n = 1 << lg
        T = gen.dna_torch(5, n, device="cuda").cpu().numpy()
        for G in Gs:
            t0 = time.time()
            rc, _, freq, st = run(lib, T, G, want_sa=False)
            print(json.dumps({"text": "dna 2^%d" % lg, "G": G, "rc": int(rc), "wall_s": round(time.time() - t0, 3), "device_s": round(st.seconds_device, 4),
                              "mbs_device": round(n / 1e6 / max(st.seconds_device, 1e-9), 1), "rounds": st.rounds, "slice_max": st.slice_max,
                              "active0": st.active_after_round0, "exchanged_GB": round(st.exchanged_bytes / 1e9, 2),
                              "verified_on_device": st.verify, "violations": st.verify_violations}), flush=True)
    print("dist64_check: %d failures" % bad)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
