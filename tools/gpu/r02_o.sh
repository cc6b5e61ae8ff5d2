#!/bin/bash
# round 2, session o: collapse with small subtrees fetched in one round trip + L2 prefetch of what children touch first + done flag on its own line
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_curves.py -m gpu -x -q --timeout 300 2>&1 | tail -4 | tee gpurun_out/r02o_tests.txt
timeout 300 python tools/trace_bench.py --tag o --rays 4194304 --check 65536 2>&1 | tail -1 | tee gpurun_out/r02o_variants.txt
LC_B200_BUILDER=lbvh timeout 300 python tools/trace_bench.py --tag o-terrain20M-lbvh --scene terrain --tris 20000000 --rays 8388608 2>&1 | tail -1 | tee -a gpurun_out/r02o_variants.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02o_build_1M_launches.csv python tools/micro/build_once.py -1000000 3 > gpurun_out/r02o_ncu1.log 2>&1
python tools/launch_summary.py gpurun_out/r02o_build_1M_launches.csv | tee gpurun_out/r02o_build_1M_launches_summary.csv
ncu --set full --clock-control none --import-source on -k regex:'k_collapse' -s 1 -c 1 -o gpurun_out/r02o_collapse -f python tools/micro/build_once.py -1000000 2 > gpurun_out/r02o_ncu3.log 2>&1; tail -1 gpurun_out/r02o_ncu3.log
