#!/bin/bash
# session r01z: verification pass after the C++ backend module / reference-pinned builtin tests — full GPU suite, smoke, default bench,
# reference arm, launch list of the bench command
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 ) 2>&1 | tee gpurun_out/r01z_gpu_tests.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r01z_bench.json 2> gpurun_out/r01z_bench.err; echo "bench rc=$?"; cat gpurun_out/r01z_bench.json; tail -3 gpurun_out/r01z_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r01z_bench_reference.json 2>/dev/null; echo "reference arm rc=$?"; cut -c1-400 gpurun_out/r01z_bench_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01z_launches.csv python bench.py --profile --steps 2 --warmup 1 > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches rc=$?"
