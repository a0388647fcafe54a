"""Tiny driver for ncu captures: one bwt_dev call on random bytes. usage: prof_child.py log2n [kind]"""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
import libsais_b200
from libsais_b200 import gen
n = 1 << int(sys.argv[1])
kind = sys.argv[2] if len(sys.argv) > 2 else "bytes"
T = gen.rand_bytes(2, n) if kind == "bytes" else (gen.dna(1, n) if kind == "dna" else gen.repetitive_dna(n // 100, 100))
ctx = libsais_b200.Context(0)
dT = torch.from_numpy(T).cuda(); dU = torch.empty(len(T), dtype=torch.uint8, device="cuda")
torch.cuda.synchronize()
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
for _ in range(reps):
    rc = ctx.bwt_dev(dT.data_ptr(), dU.data_ptr(), len(T))
print("rc", rc, ctx.stats()["device_ms"])
