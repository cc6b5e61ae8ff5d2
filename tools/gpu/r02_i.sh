#!/bin/bash
# round 2, session i: verification of the current tree (GPU suite, smoke, bench, reference arm) + launch list and ncu --set full of the timed kernel
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q --timeout 300 2>&1 | tail -8 | tee gpurun_out/r02i_tests.txt
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r02i_bench.json 2> gpurun_out/r02i_bench.err; echo "bench rc=$?"; cut -c1-600 gpurun_out/r02i_bench.json; tail -3 gpurun_out/r02i_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null > gpurun_out/r02i_bench_reference.json; cut -c1-400 gpurun_out/r02i_bench_reference.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02i_launches.csv python bench.py --profile --steps 2 --warmup 1 > gpurun_out/r02i_profile.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lc_kernel -s 1 -c 1 -o gpurun_out/r02i_c3_lc_kernel -f python bench.py --profile --steps 2 --warmup 1 > gpurun_out/r02i_ncu.log 2>&1; tail -2 gpurun_out/r02i_ncu.log
