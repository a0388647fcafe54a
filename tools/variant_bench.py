"""Experiment driver (GPU box): time the BWT pipeline per sort-pass variant / env setting.
usage: python tools/variant_bench.py [log2n] [variants...]"""
import json
import os
import subprocess
import sys

CHILD = r'''
import sys, json, numpy as np, torch
sys.path.insert(0, ".")
import libsais_b200
from libsais_b200 import gen
n = 1 << int(sys.argv[1])
kind = sys.argv[2]
T = gen.rand_bytes(2, n) if kind == "bytes" else gen.dna(1, n)
ctx = libsais_b200.Context(0)
dT = torch.from_numpy(T).cuda(); dU = torch.empty(n, dtype=torch.uint8, device="cuda")
torch.cuda.synchronize()
ctx.set_profiling(True)
for _ in range(3):
    rc = ctx.bwt_dev(dT.data_ptr(), dU.data_ptr(), n)
assert rc > 0, (rc, ctx.last_error())
st = ctx.stats()
k = st["kernels"]
sp = k["sort_pass"]
print(json.dumps({"device_ms": round(st["device_ms"], 3), "sort_pass_ms": round(sp["ms"], 3), "sort_pass_gbs": round(sp["bytes"] / sp["ms"] / 1e6, 1),
                  "launches": st["total_launches"], "kern": {a: round(b["ms"], 3) for a, b in k.items()}, "rounds": [(r["n_active"], r["passes"]) for r in st["rounds"]]}))
'''


def main():
    log2n = sys.argv[1] if len(sys.argv) > 1 else "28"
    variants = sys.argv[2:] or [str(i) for i in range(8)]
    for kind in ("bytes",):
        for v in variants:
            env = dict(os.environ, LIBSAIS_CUDA_SORT_VARIANT=v)
            r = subprocess.run([sys.executable, "-c", CHILD, log2n, kind], env=env, capture_output=True, text=True, timeout=300)
            print("variant", v, kind, r.stdout.strip() or r.stderr[-600:], flush=True)


if __name__ == "__main__":
    main()
