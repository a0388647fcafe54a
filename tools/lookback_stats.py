import sys, ctypes as C, numpy as np, torch
sys.path.insert(0, ".")
import libsais_b200
from libsais_b200 import gen
n = 1 << int(sys.argv[1])
T = gen.rand_bytes(2, n)
ctx = libsais_b200.Context(0)
dT = torch.from_numpy(T).cuda(); dU = torch.empty(n, dtype=torch.uint8, device="cuda")
out = (C.c_uint32 * 8)()
ctx.lib.libsais_cuda_debug_scalars.argtypes = [C.c_void_p, C.c_void_p, C.c_int32]
for it in range(3):
    ctx.lib.libsais_cuda_debug_scalars(ctx.handle, out, 8)
    rc = ctx.bwt_dev(dT.data_ptr(), dU.data_ptr(), n)
    ctx.lib.libsais_cuda_debug_scalars(ctx.handle, out, 8)
    w, s, r, t = out[0], out[1], out[2], out[3]
    print("tiles", t, "words/tile %.2f" % (w / max(t, 1)), "spins/tile %.2f" % (s / max(t, 1)), "refills/tile %.3f" % (r / max(t, 1)), "ms", ctx.stats()["device_ms"])
