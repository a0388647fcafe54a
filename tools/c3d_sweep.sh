#!/bin/bash
# GPU box: config 3 (full size, SA) under several settings of the position-ordered rounds; one summary line each.
mkdir -p gpurun_out
for V in "$@"; do
    env $(echo $V | tr ',' ' ') timeout 300 python tools/config_bench.py c3d > gpurun_out/sweep_$V.json 2> gpurun_out/sweep_$V.err
    python - "$V" <<'P'
import json, sys
v = sys.argv[1]
try:
    d = json.loads(open("gpurun_out/sweep_%s.json" % v).read().strip().splitlines()[0])["sa"]
    k = d["kernels"]
    print(v, "ms", d["ms"], d.get("verify"), "round", k.get("local_sort", {}).get("ms"), "apply+scatter", k.get("scatter", {}).get("ms"), "sort_pass", k.get("sort_pass", {}).get("ms"))
except Exception as e:
    print(v, "failed", e)
P
done
