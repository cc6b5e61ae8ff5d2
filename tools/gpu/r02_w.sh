#!/bin/bash
# round 2, session w: hierarchy climb with an optimistic look at the arrival flag: parity subset, build times, launch list
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_curves.py -m gpu -x -q --timeout 200 2>&1 | tail -4 | tee gpurun_out/r02w_tests.txt
timeout 200 python tools/trace_bench.py --tag w --rays 4194304 --check 65536 2>&1 | tail -1 | tee gpurun_out/r02w_variants.txt
LC_B200_BUILDER=lbvh timeout 200 python tools/trace_bench.py --tag w-terrain20M-lbvh --scene terrain --tris 20000000 --rays 8388608 2>&1 | tail -1 | tee -a gpurun_out/r02w_variants.txt
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02w_build_1M_launches.csv python tools/micro/build_once.py -1000000 3 > gpurun_out/r02w_ncu1.log 2>&1
python tools/launch_summary.py gpurun_out/r02w_build_1M_launches.csv | tee gpurun_out/r02w_build_1M_launches_summary.csv
