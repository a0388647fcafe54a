"""Bisect helper (GPU box): one text through libsais / libsais_bwt under the MSD knobs, compared with the oracle."""
import os, sys
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import _libs
from libsais_b200 import gen
rng = np.random.default_rng(77)
_ = [gen.rand_bytes(3, 8)]
rng.integers(0, 16, 200_000)
T = (rng.integers(0, 2, 100_000) + 48).astype(np.uint8)
cu, o = _libs.cuda(), _libs.oracle()
os.environ["LIBSAIS_CUDA_MSD"] = "2"
for fuse, k32 in (("0", "0"), ("1", "0"), ("0", "1"), ("1", "1")):
    os.environ["LIBSAIS_CUDA_MSD_FUSE"] = fuse; os.environ["LIBSAIS_CUDA_MSD_K32"] = k32
    rc, SA = cu.sa(T); rco, SAo = o.sa(T)
    rb, U = cu.bwt(T); rbo, Uo = o.bwt(T)
    print("fuse", fuse, "k32", k32, "sa rc", rc, "equal", bool((SA == SAo).all()), "bwt rc", rb, rbo, "equal", bool((U == Uo).all()), flush=True)
