"""Host-API timing with pageable (plain numpy) vs pinned buffers."""
import sys, time, numpy as np, torch
sys.path.insert(0, ".")
import libsais_b200
from libsais_b200 import gen
n = 1 << 28
T = gen.rand_bytes(2, n)
ctx = libsais_b200.Context(0)
U = np.empty(n, dtype=np.uint8); A = np.empty(1, dtype=np.int32)
hT = torch.from_numpy(T).pin_memory(); hU = torch.empty(n, dtype=torch.uint8).pin_memory()
for name, tp, up in (("pageable", T.ctypes.data, U.ctypes.data), ("pinned", hT.data_ptr(), hU.data_ptr())):
    for _ in range(2):
        ctx.bwt_ptr(tp, up, A.ctypes.data, n)
    t0 = time.perf_counter()
    for _ in range(5):
        rc = ctx.bwt_ptr(tp, up, A.ctypes.data, n)
    dt = (time.perf_counter() - t0) / 5
    print(name, "bwt 256 MiB: %.1f ms/call, %.1f MB/s, device %.1f ms" % (dt * 1e3, n / 1e6 / dt, ctx.stats()["device_ms"]))
SA = np.empty(n, dtype=np.int32)
for _ in range(2):
    ctx.libsais(T, SA)
t0 = time.perf_counter(); rc = ctx.libsais(T, SA); dt = time.perf_counter() - t0
print("pageable libsais 256 MiB (1 GiB SA out): %.1f ms, device %.1f ms" % (dt * 1e3, ctx.stats()["device_ms"]))
