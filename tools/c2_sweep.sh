#!/bin/bash
# GPU box: config 2 (256 MiB random bytes, BWT) under several env settings; per-kernel ms of the main kernels.
mkdir -p gpurun_out
for V in "$@"; do
    env $(echo $V | tr ',' ' ') timeout 300 python tools/config_bench.py c2 > gpurun_out/sweep2_$V.json 2> gpurun_out/sweep2_$V.err
    python - "$V" <<'P'
import json, sys
v = sys.argv[1]
try:
    d = json.loads(open("gpurun_out/sweep2_%s.json" % v).read().strip().splitlines()[0])
    b = d["bwt"]; k = b["kernels"]
    print(v, "ms", b["ms"], "roundtrip", d["unbwt"]["roundtrip"], {x: k[x]["ms"] for x in ("sort_hist", "sort_pass_gen", "part_pass", "bucket_sort", "rank_init", "bwt") if x in k})
except Exception as e:
    print(v, "failed", e)
P
done
