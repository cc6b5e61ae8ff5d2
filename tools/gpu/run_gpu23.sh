#!/bin/bash
# session 23: build arena — GPU suite, smoke, rebuild probe, C4 scene run
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 600 python tools/micro/rebuild_probe.py 2>&1 | tail -14
timeout 900 python tools/scene_bench.py --config c4 --frames 4 > gpurun_out/r01j_c4.json 2> gpurun_out/c4.err; echo "c4 rc=$?"; cat gpurun_out/r01j_c4.json; tail -3 gpurun_out/c4.err
