#!/bin/bash
# session 38: final verification pass of the round — GPU suite, smoke, default bench, launch list + full ncu capture of k_trace on the
# bench command, C2 / C4 / C5 scene runs
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | grep -v "^  File\|^Extension" | tail -4
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r01t_bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/r01t_bench.json; tail -3 gpurun_out/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01t_launches.csv python bench.py --profile --steps 2 --warmup 1 > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_trace -s 1 -c 1 -o gpurun_out/r01t_trace python bench.py --profile --steps 2 --warmup 1 > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
timeout 600 python tools/c2_bench.py > gpurun_out/r01t_c2_bench.jsonl 2> gpurun_out/c2.err; echo "c2 rc=$?"; cat gpurun_out/r01t_c2_bench.jsonl
timeout 900 python tools/scene_bench.py --config c4 --frames 4 > gpurun_out/r01t_c4.json 2> gpurun_out/c4.err; echo "c4 rc=$?"; cat gpurun_out/r01t_c4.json
timeout 900 python tools/scene_bench.py --config c5 > gpurun_out/r01t_c5.json 2> gpurun_out/c5.err; echo "c5 rc=$?"; cat gpurun_out/r01t_c5.json
LC_B200_BUILDER=lbvh timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r01t_build_1M_launches.csv python tools/micro/build_once.py -1000000 3 > /dev/null 2>&1; echo "ncu build rc=$?"
