#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu6.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu6.log
{
timeout 300 python tools/trace_bench.py --check 65536 --tag ifif
for t in 4 8; do LC_B200_TRACE_TUNE=$t timeout 300 python tools/trace_bench.py --tag ifif-$t; done
} > gpurun_out/variants6.log 2>&1
cat gpurun_out/variants6.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_trace -s 1 -c 1 -o gpurun_out/prof_trace_c python bench.py --profile --steps 2 --warmup 1 > gpurun_out/ncu_full_c.log 2>&1; echo "ncu full rc=$?"
