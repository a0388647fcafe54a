"""Timing of adversarial inputs (many rounds / extreme skew) through the device API."""
import sys, time, json
import numpy as np, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import libsais_b200, cases
from libsais_b200 import gen
ctx = libsais_b200.Context(0)
def run(name, T):
    n = len(T)
    dT = torch.from_numpy(T).cuda(); dSA = torch.empty(n, dtype=torch.int32, device="cuda")
    t0 = time.time(); rc = ctx.sa_dev(dT.data_ptr(), dSA.data_ptr(), n); dt = time.time() - t0
    st = ctx.stats()
    print(json.dumps({"case": name, "n": n, "rc": rc, "wall_ms": round(dt * 1e3, 1), "device_ms": round(st["device_ms"], 1),
                      "rounds": len(st["rounds"]), "launches": st["total_launches"], "mbs": round(n / 1e6 / (st["device_ms"] / 1e3), 1)}), flush=True)
log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 26
n = 1 << log2n
run("zeros", np.zeros(n, dtype=np.uint8))
run("abab", np.resize(np.frombuffer(b"ab", dtype=np.uint8), n))
run("fib", cases.fibonacci_string(37)[:n].copy())
run("thue_morse", cases.thue_morse(log2n))
run("period_1000", np.resize(gen.dna(3, 1000), n))
run("two_copies", np.concatenate([gen.dna(4, n // 2), gen.dna(4, n // 2)]))
run("english_like", (np.random.default_rng(1).integers(0, 27, n) + 96).astype(np.uint8))
