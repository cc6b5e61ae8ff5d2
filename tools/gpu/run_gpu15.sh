#!/bin/bash
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_collapse|k_hierarchy" -s 2 -c 2 -o gpurun_out/prof_build python tools/trace_bench.py --rays 262144 --reps 1 > gpurun_out/ncu_build_full.log 2>&1; echo "ncu rc=$?"
