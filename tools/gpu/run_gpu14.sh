#!/bin/bash
timeout 300 python tools/trace_bench.py --rays 1048576 --reps 2 --check 20000 --tag poll 2>&1 | tail -1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_build.csv python tools/trace_bench.py --rays 262144 --reps 1 > gpurun_out/ncu_build.log 2>&1; echo "ncu rc=$?"
