#!/bin/bash
# Run on the GPU box (gpurun): the full -m gpu suite, the default bench line, the ncu launch list of one bench step and
# (optionally) the per-kernel ncu captures, SUMMARISED ON THE BOX -- the .ncu-rep files are too large for gpurun_out/.
#   usage: bash tools/gpu_check.sh <tag> [tests] [bench] [launches] [prof_c2] [prof_c3s]
set -u
mkdir -p gpurun_out
R=${1:-r2}; shift
WHAT=" ${*:-tests bench} "
if [[ "$WHAT" == *" tests "* ]]; then
    timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$R.log 2>&1
    echo "pytest rc $? : $(tail -n 1 gpurun_out/pytest_$R.log)"
fi
if [[ "$WHAT" == *" bench "* ]]; then
    timeout 900 python bench.py > gpurun_out/bench_$R.json 2> gpurun_out/bench_$R.err
    echo "bench rc $?"
fi
if [[ "$WHAT" == *" launches "* ]]; then
    timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$R.csv \
        python tools/prof_child.py 28 bytes 3 > gpurun_out/launches_$R.log 2>&1
    echo "launch list rc $?"
fi
for W in c2 c3s; do
    if [[ "$WHAT" == *" prof_$W "* ]]; then
        timeout 1200 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -f -o /tmp/prof_${W}_$R \
            python tools/prof_kernels.py $W > gpurun_out/prof_${W}_$R.log 2>&1
        echo "capture $W rc $?"
        python tools/ncu_table.py /tmp/prof_${W}_$R.ncu-rep > gpurun_out/kernels_${W}_$R.md 2>> gpurun_out/prof_${W}_$R.log
        ls -l /tmp/prof_${W}_$R.ncu-rep
    fi
done
