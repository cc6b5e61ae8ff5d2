#!/bin/bash
# session 25: ncu --set full of the two latency-bound build kernels on the C4-sized terrain and on the C3 soup
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_collapse|k_hierarchy" -c 2 -o gpurun_out/r01l_build_c4 python tools/micro/build_once.py 3164 > gpurun_out/ncu_build_c4.log 2>&1; echo "ncu c4 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_collapse|k_hierarchy" -c 2 -o gpurun_out/r01l_build_c3 python tools/micro/build_once.py -1000000 > gpurun_out/ncu_build_c3.log 2>&1; echo "ncu c3 rc=$?"
ls -la gpurun_out/*.ncu-rep
