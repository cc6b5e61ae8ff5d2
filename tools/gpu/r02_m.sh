#!/bin/bash
# round 2, session m: collapse on full-warp butterflies (no partial-mask redux), compact slot->primitive list, lighter fences, 32-bit sort keys with
# small tiles, 8-lane refit: full GPU suite, build / refit times at 1 M and 20 M, launch lists
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q --timeout 300 2>&1 | tail -6 | tee gpurun_out/r02m_tests.txt
timeout 300 python tools/trace_bench.py --tag m --check 65536 2>&1 | tail -1 | tee gpurun_out/r02m_build.txt
LC_B200_BUILDER=lbvh timeout 300 python tools/trace_bench.py --tag m-terrain20M-lbvh --scene terrain --tris 20000000 --rays 8388608 2>&1 | tail -1 | tee -a gpurun_out/r02m_build.txt
timeout 300 python tools/trace_bench.py --tag m-terrain20M-auto --scene terrain --tris 20000000 --rays 8388608 2>&1 | tail -1 | tee -a gpurun_out/r02m_build.txt
timeout 300 python tools/micro/rebuild_probe.py 2>&1 | tail -12 | tee gpurun_out/r02m_rebuild_probe.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02m_build_1M_launches.csv python tools/micro/build_once.py -1000000 3 > gpurun_out/r02m_ncu1.log 2>&1
python tools/launch_summary.py gpurun_out/r02m_build_1M_launches.csv | tee gpurun_out/r02m_build_1M_launches_summary.csv
LC_B200_BUILDER=lbvh ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02m_build_20M_launches.csv python tools/micro/build_once.py 3164 2 > gpurun_out/r02m_ncu2.log 2>&1
python tools/launch_summary.py gpurun_out/r02m_build_20M_launches.csv | tee gpurun_out/r02m_build_20M_launches_summary.csv
ncu --set full --clock-control none --import-source on -k regex:'k_collapse|k_hierarchy' -s 2 -c 2 -o gpurun_out/r02m_build_kernels -f python tools/micro/build_once.py -1000000 2 > gpurun_out/r02m_ncu3.log 2>&1; tail -2 gpurun_out/r02m_ncu3.log
