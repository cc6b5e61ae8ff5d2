#!/bin/bash
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu13.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu13.log
timeout 300 python tools/trace_bench.py --check 65536 --tag newbuild 2>&1 | tail -2
timeout 600 python tools/scene_bench.py --config c4 --frames 2 2>&1 | tail -2
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_build.csv python tools/trace_bench.py --rays 262144 --reps 1 > gpurun_out/ncu_build.log 2>&1; echo "ncu rc=$?"
