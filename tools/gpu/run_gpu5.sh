#!/bin/bash
mkdir -p gpurun_out
{
timeout 300 python tools/trace_bench.py --check 65536 --tag ww-default
for t in 8,8,32,4 8,8,32,8 8,8,32,12 8,8,32,16 6,8,32,0 10,8,32,0 6,8,32,8 8,8,32,6; do
LC_B200_TRACE_TUNE=$t timeout 300 python tools/trace_bench.py --tag ww-$t
done
} > gpurun_out/variants5.log 2>&1
cat gpurun_out/variants5.log
