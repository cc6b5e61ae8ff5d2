#!/bin/bash
timeout 1200 python -m pytest tests/test_parity_gpu.py -m gpu -x -q 2>&1 | grep -v "^  File\|^Extension" | tail -4
LC_B200_BUILDER=lbvh timeout 300 python tools/trace_bench.py --tag lbvh_warp --check 20000 2>&1 | tail -1
LC_B200_BUILDER=lbvh timeout 300 python tools/trace_bench.py --tag lbvh_warp --scene terrain --tris 20000000 --rays 8388608 2>&1 | tail -1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_hierarchy -c 3 python tools/micro/build_once.py -1000000 3 2>&1 | grep -i "k_hierarchy\|gpu__time" | tail -6
