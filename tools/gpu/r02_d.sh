#!/bin/bash
# round 2, session d: full GPU suite after the async build / download changes, k_trace A/B (identity-instance shortcut, CTAs per SM), C2 and C5 with direct / wavefront
# lowering x precise / fast math, then the bench line
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q --timeout 300 2>&1 | tail -12 | tee gpurun_out/r02d_tests.txt
for v in default noident mb6; do
  lib=""; [ $v != default ] && lib=luisa-compute-rs_b200/lib/variants/liblc_b200_$v.so
  LC_B200_LIB=$lib timeout 300 python tools/trace_bench.py --tag $v 2>&1 | tail -1
done | tee gpurun_out/r02d_k_trace_variants.txt
timeout 400 python tools/dsl_bench.py c2 --configs 1:4:8:0,1:4:8:1,0:4:32:0,0:4:32:1,0:4:16:1 2>/dev/null | tee gpurun_out/r02d_dsl_c2.jsonl
for low in auto direct; do for fm in "" "--precise"; do
  timeout 300 python tools/c5_path_trace.py --spp 64 --lowering $low $fm 2>/dev/null | tail -1
done; done | tee gpurun_out/r02d_c5_n1.jsonl
timeout 900 python bench.py > gpurun_out/r02d_bench.json 2> gpurun_out/r02d_bench.err; echo "bench rc=$?"; cut -c1-1500 gpurun_out/r02d_bench.json; tail -3 gpurun_out/r02d_bench.err
