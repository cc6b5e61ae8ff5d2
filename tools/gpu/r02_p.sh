#!/bin/bash
# round 2, session p: slot-assignment shortcut in the collapse; full GPU suite; build launch list
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q --timeout 300 2>&1 | tail -4 | tee gpurun_out/r02p_tests.txt
timeout 300 python tools/trace_bench.py --tag p --rays 4194304 --check 65536 2>&1 | tail -1 | tee gpurun_out/r02p_variants.txt
LC_B200_BUILDER=lbvh timeout 300 python tools/trace_bench.py --tag p-terrain20M-lbvh --scene terrain --tris 20000000 --rays 8388608 2>&1 | tail -1 | tee -a gpurun_out/r02p_variants.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02p_build_1M_launches.csv python tools/micro/build_once.py -1000000 3 > gpurun_out/r02p_ncu1.log 2>&1
python tools/launch_summary.py gpurun_out/r02p_build_1M_launches.csv | tee gpurun_out/r02p_build_1M_launches_summary.csv
