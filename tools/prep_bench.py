import sys, time, ctypes as C, torch
sys.path.insert(0, ".")
import libsais_b200
from libsais_b200 import gen
from libsais_b200.dist import _Lib
L = _Lib().lib
ctx = libsais_b200.Context(0)
for n in [(1 << 30), (1 << 30) - 32, (1 << 29), (1 << 31)]:
    dT = gen.dna_torch(5, n, device="cuda")
    torch.cuda.synchronize()
    k, kb = C.c_int32(0), C.c_int32(0)
    for it in range(3):
        t0 = time.perf_counter()
        rc = L.libsais_cuda_dist_prepare(ctx.handle, dT.data_ptr(), n, C.byref(k), C.byref(kb))
        dt = time.perf_counter() - t0
        st = ctx.stats()
        print(n, "iter", it, "rc", rc, "prepare ms %.2f" % (dt * 1e3), {a: round(b["ms"], 3) for a, b in st["kernels"].items()}, flush=True)
    del dT
