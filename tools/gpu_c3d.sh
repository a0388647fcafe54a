#!/bin/bash
# GPU box: launch list of config 3 (full size, SA) + full captures of the round kernels, summarised on the box.
set -u
R=$1
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c3d_$R.csv python tools/prof_kernels.py c3d > /tmp/l.log 2>&1
echo "launch list rc $?"


bash tools/prof_each.sh $R c3d 'po_round_kernel<\(bool\)1>' 'po_round_kernel<\(bool\)0>' 'po_apply_kernel'
