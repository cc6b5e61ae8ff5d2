#!/bin/bash
# session 22: smoke (with the lowered kernel), default bench, full-size C4 / C5 scene runs, launch list of the 1M / 20M builds
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r01i_bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/r01i_bench.json; tail -3 gpurun_out/bench.err
timeout 900 python tools/scene_bench.py --config c4 --frames 4 > gpurun_out/r01i_c4.json 2> gpurun_out/c4.err; echo "c4 rc=$?"; cat gpurun_out/r01i_c4.json; tail -3 gpurun_out/c4.err
timeout 900 python tools/scene_bench.py --config c5 > gpurun_out/r01i_c5.json 2> gpurun_out/c5.err; echo "c5 rc=$?"; cat gpurun_out/r01i_c5.json; tail -3 gpurun_out/c5.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r01i_launches_c4.csv python tools/scene_bench.py --config c4 --frames 1 > gpurun_out/ncu_c4.log 2>&1; echo "ncu rc=$?"
