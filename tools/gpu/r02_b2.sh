#!/bin/bash
# round 2: grid-stride box kernels + smaller queue clear (variant `boxes`): parity subset, build times at 1 M / 20 M, launch list
mkdir -p gpurun_out
export LC_B200_LIB=luisa-compute-rs_b200/lib/variants/liblc_b200_boxes.so
timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_curves.py -m gpu -x -q --timeout 200 2>&1 | tail -3 | tee gpurun_out/r02b2_tests.txt
timeout 200 python tools/trace_bench.py --tag boxes --rays 4194304 --check 65536 2>&1 | tail -1 | tee gpurun_out/r02b2_build.txt
LC_B200_BUILDER=lbvh timeout 200 python tools/trace_bench.py --tag boxes-terrain20M-lbvh --scene terrain --tris 20000000 --rays 8388608 2>&1 | tail -1 | tee -a gpurun_out/r02b2_build.txt
LC_B200_BUILDER=lbvh timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file gpurun_out/r02b2_build_20M_launches.csv python tools/micro/build_once.py 3164 2 > gpurun_out/r02b2_ncu.log 2>&1
python tools/launch_summary.py gpurun_out/r02b2_build_20M_launches.csv | tee gpurun_out/r02b2_build_20M_launches_summary.csv
