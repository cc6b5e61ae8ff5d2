#!/bin/bash
# round 2, session b: ncu --set full of the wavefront-lowered kernels (C2 path tracer, C3 trace kernel) next to the direct C2 kernel
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on -k regex:lc_kernel"
$NCU -s 3 -c 1 -o gpurun_out/r02b_c2_wave -f timeout 600 python tools/dsl_bench.py c2 --configs 0:4:32 > gpurun_out/r02b_c2_wave.log 2>&1; tail -2 gpurun_out/r02b_c2_wave.log
$NCU -s 3 -c 1 -o gpurun_out/r02b_c2_wave_y8 -f timeout 600 python tools/dsl_bench.py c2 --configs 0:4:8 > gpurun_out/r02b_c2_wave_y8.log 2>&1; tail -2 gpurun_out/r02b_c2_wave_y8.log
$NCU -s 3 -c 1 -o gpurun_out/r02b_c3_wave -f timeout 600 python tools/dsl_bench.py c3 --configs 0:5:8 > gpurun_out/r02b_c3_wave.log 2>&1; tail -2 gpurun_out/r02b_c3_wave.log
ls -la gpurun_out/*.ncu-rep
