#!/bin/bash
# Run on the GPU box: one `ncu --set full` capture of the FIRST launch of every kernel named below (one ncu run per
# kernel: the workload is short, and a capture of all launches of config 3 does not finish in the box's time limit),
# summarised on the box into gpurun_out/kernels_<workload>_<tag>.md (+ the stall hot spots per kernel in
# gpurun_out/hot_<kernel>_<tag>.txt).  The .ncu-rep files stay in /tmp: they are too large for gpurun_out/.
#   usage: bash tools/prof_each.sh <tag> <c2|c3s> kernel_regex [kernel_regex ...]
set -u
R=$1; W=$2; shift 2
mkdir -p gpurun_out
REPS=""
for K in "$@"; do
    F=/tmp/prof_${W}_$(echo "$K" | tr -c 'A-Za-z0-9_' '_')
    timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$K" -c 1 -f -o $F \
        python tools/prof_kernels.py $W > /tmp/prof_each.log 2>&1
    echo "capture $K rc $?"
    if [ -f $F.ncu-rep ]; then
        REPS="$REPS $F.ncu-rep"
        python tools/ncu_summary.py $F.ncu-rep 14 > gpurun_out/hot_$(basename $F)_$R.txt 2>&1
    fi
done
python tools/ncu_table.py $REPS > gpurun_out/kernels_${W}_$R.md
