#!/bin/bash
# C++ host check on B200 + initcheck after zeroing the whole build header
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_cpp_backend.py -q 2>&1 | tail -5
( cd oracle/_ref/cpp_host && timeout 120 ./lc_cpp_host_check ) > gpurun_out/r01z_cpp_host_check.log 2>&1; echo "host rc=$?"; cat gpurun_out/r01z_cpp_host_check.log | cut -c1-250 | tail -12
CS=/usr/local/cuda/bin/compute-sanitizer
K="c1_triangle or c2_cornell_primary or instances_transforms or empty_inputs or degenerate or duplicate_centroids or rebuild_after or ray_query or ties_lowest or float3_stride or prefer_update"
( time timeout 300 $CS --tool initcheck --print-limit 3000 --error-exitcode 9 python -m pytest tests/test_parity_gpu.py -m gpu -q -k "$K" ) > gpurun_out/r01z_initcheck_tests2.log 2>&1; echo "initcheck rc=$?"
grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r01z_initcheck_tests2.log
grep -E "Uninitialized|Host API memory access" gpurun_out/r01z_initcheck_tests2.log | sed -E 's/0x[0-9a-f]+/X/g' | sort | uniq -c | sort -rn | head
