"""Is the onesweep pass sensitive to power-of-two sizes (DRAM partition camping)?"""
import sys, json, torch
sys.path.insert(0, ".")
import libsais_b200
from libsais_b200 import gen
ctx = libsais_b200.Context(0); ctx.set_profiling(True)
for n in [(1 << 29), (1 << 29) + 1234567, (1 << 28), (1 << 28) - 7777, (1 << 30), (1 << 30) - 99991]:
    dT = gen.dna_torch(5, n, device="cuda")
    dSA = torch.empty(n, dtype=torch.int32, device="cuda")
    for _ in range(2):
        rc = ctx.sa_dev(dT.data_ptr(), dSA.data_ptr(), n)
    st = ctx.stats(); k = st["kernels"]
    sp = k["sort_pass"]
    print(json.dumps({"n": n, "rc": rc, "device_ms": round(st["device_ms"], 2), "mbs": round(n / 1e6 / (st["device_ms"] / 1e3)),
                      "sort_pass_ms": round(sp["ms"], 2), "sort_pass_gbs": round(sp["bytes"] / sp["ms"] / 1e6), "passes": sp["launches"],
                      "rank": round(k["rank_init"]["ms"], 2), "gen": round(k["sort_pass_gen"]["ms"], 2)}), flush=True)
    del dT, dSA
