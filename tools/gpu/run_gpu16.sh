#!/bin/bash
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_c4.csv python tools/scene_bench.py --config c4 --frames 1 > gpurun_out/ncu_c4.log 2>&1; echo "ncu rc=$?"
