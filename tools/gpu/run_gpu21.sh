#!/bin/bash
# session 21: IR -> CUDA lowering on the GPU — full GPU suite, C2 through create_shader vs the hand-lowered kernel
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 600 python tools/c2_bench.py --size 1024 --spp 32 --dispatches 8 > gpurun_out/r01h_c2_bench.jsonl 2> gpurun_out/c2_bench.err; echo "c2 rc=$?"; cat gpurun_out/r01h_c2_bench.jsonl; tail -5 gpurun_out/c2_bench.err
