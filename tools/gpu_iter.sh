#!/bin/bash
# One development iteration on the GPU box: the round-2 path tests, a short differential stress, config benches.
#   usage: bash tools/gpu_iter.sh <tag> [config ...]
set -u
R=$1; shift
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_round2.py tests/test_cuda_parity.py -m gpu -x -q > gpurun_out/it_pytest_$R.log 2>&1
echo "pytest rc $? : $(tail -n 1 gpurun_out/it_pytest_$R.log)"
timeout 200 python tools/stress.py 45 21 > gpurun_out/it_stress_$R.log 2>&1
echo "stress rc $? : $(tail -n 1 gpurun_out/it_stress_$R.log)"
timeout 200 python tools/stress.py 45 22 big >> gpurun_out/it_stress_$R.log 2>&1
echo "stress big rc $? : $(tail -n 1 gpurun_out/it_stress_$R.log)"
for C in "$@"; do
    timeout 600 python tools/config_bench.py $C > gpurun_out/it_${C}_$R.json 2> gpurun_out/it_${C}_$R.err
    echo "$C rc $? $(head -c 300 gpurun_out/it_${C}_$R.json)"
done
