#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_trace -s 1 -c 1 -o gpurun_out/prof_trace_b python bench.py --profile --steps 2 --warmup 1 > gpurun_out/ncu_full_b.log 2>&1; echo "ncu full rc=$?"
