#!/bin/bash
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu9.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu9.log
{
timeout 300 python tools/trace_bench.py --check 65536 --tag sort-default
for t in 0 1,262144,5,2 1,262144,6,2 1,262144,4,3 1,262144,7,2 1,262144,6,3 1,262144,5,4 1,262144,8,1; do
LC_B200_RAY_SORT=$t timeout 300 python tools/trace_bench.py --tag sort-$t
done
} > gpurun_out/variants9.log 2>&1
cat gpurun_out/variants9.log
