#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu2.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu2.log
V=luisa-compute-rs_b200/lib/variants
{
timeout 300 python tools/trace_bench.py --check 65536 --tag mb5
LC_B200_LIB=$V/liblc_b200_mb6.so timeout 300 python tools/trace_bench.py --tag mb6
LC_B200_LIB=$V/liblc_b200_mb4.so timeout 300 python tools/trace_bench.py --tag mb4
LC_B200_PHASE_WEIGHTS=1,1,1,1 LC_B200_LIB=$V/liblc_b200_mb6.so timeout 300 python tools/trace_bench.py --tag mb6-w1111
LC_B200_PHASE_WEIGHTS=2,3,3,4 LC_B200_LIB=$V/liblc_b200_mb6.so timeout 300 python tools/trace_bench.py --tag mb6-w2334
LC_B200_PHASE_WEIGHTS=3,2,2,4 LC_B200_LIB=$V/liblc_b200_mb6.so timeout 300 python tools/trace_bench.py --tag mb6-w3224
LC_B200_LIB=$V/liblc_b200_mb6.so timeout 300 python tools/trace_bench.py --tag mb6-terrain --scene terrain --tris 2000000 --check 65536
} > gpurun_out/variants2.log 2>&1
cat gpurun_out/variants2.log
