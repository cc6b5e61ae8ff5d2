#!/bin/bash
# compute-sanitizer passes over the small parity cases (memcheck, racecheck on the shared-memory kernels, synccheck)
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
K="c1_triangle or c2_cornell_primary or instances_transforms or empty_inputs or degenerate or duplicate_centroids or rebuild_after or ray_query or ties_lowest or float3_stride"
( time timeout 200 $CS --tool memcheck --error-exitcode 9 python __graft_entry__.py smoke ) > gpurun_out/r01z_memcheck_smoke.log 2>&1; echo "memcheck smoke rc=$?"
( time timeout 260 $CS --tool memcheck --error-exitcode 9 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "$K" ) > gpurun_out/r01z_memcheck_tests.log 2>&1; echo "memcheck tests rc=$?"
( time timeout 200 $CS --tool racecheck --error-exitcode 9 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "c1_triangle or instances_transforms or duplicate_centroids" ) > gpurun_out/r01z_racecheck_tests.log 2>&1; echo "racecheck tests rc=$?"
tail -4 gpurun_out/r01z_memcheck_smoke.log gpurun_out/r01z_memcheck_tests.log gpurun_out/r01z_racecheck_tests.log
grep -c "ERROR SUMMARY" gpurun_out/r01z_*.log
