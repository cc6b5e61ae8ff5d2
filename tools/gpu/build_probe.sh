#!/bin/bash
# builder changes on one B200 (LIB=variant name under lib/variants, default: the main library; TAG=name of the output files):
# parity subset, build + trace times on the 1 M soup and the 20 M terrain (LBVH and the default builder), refit, ncu launch lists of a 1 M and a 20 M build.
# A kernel that can spin (work queues, look-back) goes FIRST through `timeout 60 python tools/micro/build_once.py -100000 1` on its own.
mkdir -p gpurun_out
T=${TAG:-build}
[ -n "$LIB" ] && export LC_B200_LIB=luisa-compute-rs_b200/lib/variants/liblc_b200_$LIB.so
timeout 60 python tools/micro/build_once.py -100000 1 | tail -1 || { echo "small build failed or hung: stopping"; exit 1; }
timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_curves.py -m gpu -x -q --timeout 200 2>&1 | tail -3 | tee gpurun_out/${T}_tests.txt
timeout 200 python tools/trace_bench.py --tag $T --rays 4194304 --check 65536 2>&1 | tail -1 | tee gpurun_out/${T}_build.txt
LC_B200_BUILDER=lbvh timeout 200 python tools/trace_bench.py --tag $T-terrain20M-lbvh --scene terrain --tris 20000000 --rays 8388608 2>&1 | tail -1 | tee -a gpurun_out/${T}_build.txt
timeout 200 python tools/trace_bench.py --tag $T-terrain20M-auto --scene terrain --tris 20000000 --rays 8388608 2>&1 | tail -1 | tee -a gpurun_out/${T}_build.txt
timeout 200 python tools/micro/rebuild_probe.py 2>&1 | sed -n 2,6p | tee -a gpurun_out/${T}_build.txt
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${T}_1M_launches.csv python tools/micro/build_once.py -1000000 3 > gpurun_out/${T}_ncu1.log 2>&1
python tools/launch_summary.py gpurun_out/${T}_1M_launches.csv | tee gpurun_out/${T}_1M_launches_summary.csv
LC_B200_BUILDER=lbvh timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${T}_20M_launches.csv python tools/micro/build_once.py 3164 2 > gpurun_out/${T}_ncu2.log 2>&1
python tools/launch_summary.py gpurun_out/${T}_20M_launches.csv | tee gpurun_out/${T}_20M_launches_summary.csv
