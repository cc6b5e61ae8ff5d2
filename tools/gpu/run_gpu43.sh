#!/bin/bash
# wider compute-sanitizer passes: memcheck over every GPU test except the full-size ones, racecheck on curves / IR kernels, initcheck on the small parity cases
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
( time timeout 900 $CS --tool memcheck --error-exitcode 9 python -m pytest tests -m gpu -q -k "not full_size" ) > gpurun_out/r01z_memcheck_all.log 2>&1; echo "memcheck all rc=$?"
( time timeout 300 $CS --tool racecheck --error-exitcode 9 python -m pytest tests/test_curves.py tests/test_ir_lowering.py -m gpu -q -x ) > gpurun_out/r01z_racecheck_ir_curves.log 2>&1; echo "racecheck ir/curves rc=$?"
K="c1_triangle or c2_cornell_primary or instances_transforms or empty_inputs or degenerate or duplicate_centroids or rebuild_after or ray_query or ties_lowest or float3_stride or prefer_update"
( time timeout 300 $CS --tool initcheck --error-exitcode 9 python -m pytest tests/test_parity_gpu.py -m gpu -q -k "$K" ) > gpurun_out/r01z_initcheck_tests.log 2>&1; echo "initcheck rc=$?"
for f in gpurun_out/r01z_memcheck_all.log gpurun_out/r01z_racecheck_ir_curves.log gpurun_out/r01z_initcheck_tests.log; do echo "== $f"; tail -n 8 $f; done
