"""PLCP/LCP timing and parity on adversarial inputs (long matches) through the device API; the compiled
reference gives the expected arrays.  usage: python tools/plcp_patho.py [log2n]"""
import sys, time, json
import numpy as np, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import libsais_b200, _libs
from libsais_b200 import gen
ctx = libsais_b200.Context(0)
ref = _libs.ref() or _libs.oracle()
def run(name, T):
    n = len(T)
    dT = torch.from_numpy(T).cuda(); dSA = torch.empty(n, dtype=torch.int32, device="cuda")
    dP = torch.empty(n, dtype=torch.int32, device="cuda"); dL = torch.empty(n, dtype=torch.int32, device="cuda")
    rc = ctx.sa_dev(dT.data_ptr(), dSA.data_ptr(), n); sa_ms = ctx.stats()["device_ms"]
    t0 = time.time(); rc2 = ctx.plcp_dev(dT.data_ptr(), dSA.data_ptr(), dP.data_ptr(), n); wall = time.time() - t0
    st = ctx.stats()
    rc3 = ctx.lcp_dev(dP.data_ptr(), dSA.data_ptr(), dL.data_ptr(), n); lcp_ms = ctx.stats()["device_ms"]
    SA = dSA.cpu().numpy()
    r, P = ref.plcp(T, SA); r2, L = ref.lcp(P, SA)
    ok = bool((dP.cpu().numpy() == P).all()) and bool((dL.cpu().numpy() == L).all())
    print(json.dumps({"case": name, "n": n, "rc": [rc, rc2, rc3], "parity": ok, "sa_ms": round(sa_ms, 2), "plcp_ms": round(st["device_ms"], 2),
                      "plcp_wall_ms": round(wall * 1e3, 1), "lcp_ms": round(lcp_ms, 2), "launches": st["total_launches"]}), flush=True)
log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
n = 1 << log2n
run("random", gen.rand_bytes(2, n))
run("zeros", np.zeros(n, dtype=np.uint8))
run("abab", np.resize(np.frombuffer(b"ab", dtype=np.uint8), n))
run("period_1000", np.resize(gen.dna(3, 1000), n))
run("two_copies", np.concatenate([gen.dna(4, n // 2), gen.dna(4, n // 2)]))
run("c3_scaled", gen.repetitive_dna(n // 100, 100))
