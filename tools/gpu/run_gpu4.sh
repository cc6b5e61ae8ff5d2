#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu4.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu4.log
V=luisa-compute-rs_b200/lib/variants
{
LC_B200_LIB=$V/liblc_b200_vote.so timeout 300 python tools/trace_bench.py --tag vote
timeout 300 python tools/trace_bench.py --check 65536 --tag ww-8,2,4
for t in 4,2,4 12,2,4 16,2,4 8,0,0 8,4,8 8,8,32 8,3,6 2,2,4; do
LC_B200_TRACE_TUNE=$t timeout 300 python tools/trace_bench.py --tag ww-$t
done
} > gpurun_out/variants4.log 2>&1
cat gpurun_out/variants4.log
