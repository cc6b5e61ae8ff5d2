#!/bin/bash
V=luisa-compute-rs_b200/lib/variants
{
for v in mb5 mb7 mb8; do LC_B200_LIB=$V/liblc_b200_$v.so timeout 300 python tools/trace_bench.py --tag $v; done
} > gpurun_out/variants8.log 2>&1
cat gpurun_out/variants8.log
