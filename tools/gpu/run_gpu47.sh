#!/bin/bash
# CTA size of the persistent traversal kernel (warps are independent): 128 x 7 (default) vs 64 x 14 vs 32 x 28 per SM — kernel alone and host pipeline
mkdir -p gpurun_out
{
for v in default t64 t32; do
  lib=luisa-compute-rs_b200/lib/liblc_b200.so; [ $v != default ] && lib=luisa-compute-rs_b200/lib/variants/liblc_b200_$v.so
  echo "== $v"
  LC_B200_LIB=$PWD/$lib timeout 300 python tools/trace_bench.py --reps 5 --tag $v 2>/dev/null | tail -2
  LC_B200_LIB=$PWD/$lib timeout 200 python tools/e2e_bench.py 2>/dev/null | tail -1
done
} | tee gpurun_out/r01z_cta_size.txt
