#!/bin/bash
# session 20 (re-entry): verify restored HEAD, launch list of the default bench, full ncu capture of k_trace and of the build
mkdir -p gpurun_out
bash tools/gpu/verify.sh
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01g_launches.csv python bench.py --profile --steps 2 --warmup 1 > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_trace -s 1 -c 1 -o gpurun_out/r01g_trace python bench.py --profile --steps 2 --warmup 1 > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out
