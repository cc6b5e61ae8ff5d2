#!/bin/bash
# memcheck + racecheck over the tests added in session r01z (reference-pinned builtins / texels / sampling, cut-out path tracer, C++ host)
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
( time timeout 600 $CS --tool memcheck --error-exitcode 9 python -m pytest tests/test_device_math_reference.py tests/test_ir_lowering.py -m gpu -q -k "device_math or texel or texture3d or builtins or reverse or casts or cutout" ) > gpurun_out/r01z_memcheck_new.log 2>&1; echo "memcheck rc=$?"
( time timeout 600 $CS --tool racecheck --error-exitcode 9 python -m pytest tests/test_device_math_reference.py tests/test_ir_lowering.py -m gpu -q -k "texel or texture3d or cutout" ) > gpurun_out/r01z_racecheck_new.log 2>&1; echo "racecheck rc=$?"
( cd oracle/_ref/cpp_host && timeout 200 $CS --tool memcheck --error-exitcode 9 ./lc_cpp_host_check ) > gpurun_out/r01z_memcheck_cpp_host.log 2>&1; echo "memcheck cpp host rc=$?"
grep -E "passed|failed|ERROR SUMMARY|RACECHECK SUMMARY|cpp_host_check ok" gpurun_out/r01z_memcheck_new.log gpurun_out/r01z_racecheck_new.log gpurun_out/r01z_memcheck_cpp_host.log
