#!/bin/bash
# round 2, session l: barrier-free collapse (ticket queue + child preloading) and batched look-back in the onesweep: parity subset, build times
# at 1 M / 20 M, launch list of a 1 M build, ncu --set full of the three build kernels that dominate
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_curves.py -m gpu -x -q --timeout 300 2>&1 | tail -6 | tee gpurun_out/r02l_tests.txt
timeout 300 python tools/trace_bench.py --tag collapse2 --check 65536 2>&1 | tail -1 | tee gpurun_out/r02l_build.txt
timeout 300 python tools/trace_bench.py --tag collapse2-terrain20M --scene terrain --tris 20000000 --rays 8388608 2>&1 | tail -1 | tee -a gpurun_out/r02l_build.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02l_build_1M_launches.csv python tools/micro/build_once.py -1000000 3 > gpurun_out/r02l_ncu1.log 2>&1
python tools/launch_summary.py gpurun_out/r02l_build_1M_launches.csv | tee gpurun_out/r02l_build_1M_launches_summary.csv
ncu --set full --clock-control none --import-source on -k regex:'k_collapse|k_hierarchy|k_sort_onesweep' -s 5 -c 6 -o gpurun_out/r02l_build_kernels -f python tools/micro/build_once.py -1000000 2 > gpurun_out/r02l_ncu2.log 2>&1; tail -2 gpurun_out/r02l_ncu2.log
