#!/bin/bash
# Run on the GPU box (gpurun): regenerates the raw material of profiles/ into gpurun_out/.
#   1. per-launch device times of one bench step (ncu, cold-cache, serialised: compare SHARES)
#   2. one full capture (--set full) of a full-size onesweep digit pass
#   3. the bench line itself (not under a profiler)
set -u
mkdir -p gpurun_out
R=${1:-r1}
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$R.json 2> gpurun_out/bench_$R.err
echo "bench rc $?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 81 -c 60 --csv --log-file gpurun_out/launches_$R.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/launches_$R.log 2>&1
echo "launch list rc $?"
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:sort_pass_kernel.*NoGen' -s 30 -c 1 -f -o gpurun_out/prof_sort_pass_$R \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/prof_sort_pass_$R.log 2>&1
echo "full capture rc $?"
