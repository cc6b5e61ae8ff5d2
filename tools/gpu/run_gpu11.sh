#!/bin/bash
timeout 900 python tools/scene_bench.py --config c4 --nx 1001 --check 20000 > gpurun_out/c4_small.json 2> gpurun_out/c4_small.err; echo rc=$?; cat gpurun_out/c4_small.json; tail -3 gpurun_out/c4_small.err
timeout 900 python tools/scene_bench.py --config c4 > gpurun_out/c4_full.json 2> gpurun_out/c4_full.err; echo rc=$?; cat gpurun_out/c4_full.json; tail -3 gpurun_out/c4_full.err
timeout 900 python tools/scene_bench.py --config c5 > gpurun_out/c5_full.json 2> gpurun_out/c5_full.err; echo rc=$?; cat gpurun_out/c5_full.json; tail -3 gpurun_out/c5_full.err
