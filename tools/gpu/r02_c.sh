#!/bin/bash
# round 2, session c: full GPU test suite, smoke, the new bench line (DSL headline, e2e through DeviceInterface, C5 leg), C2 after the identity-instance shortcut,
# launch list + ncu --set full of the timed kernel
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/r02c_tests.txt
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r02c_bench.json 2> gpurun_out/r02c_bench.err; echo "bench rc=$?"; cat gpurun_out/r02c_bench.json; tail -5 gpurun_out/r02c_bench.err
timeout 300 python tools/dsl_bench.py c2 --configs 1:4:8,0:4:32,0:4:16 2>/dev/null | tee gpurun_out/r02c_dsl_c2.jsonl
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02c_launches.csv python bench.py --profile --steps 2 --warmup 1 > gpurun_out/r02c_profile.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lc_kernel -s 1 -c 1 -o gpurun_out/r02c_c3_lc_kernel -f python bench.py --profile --steps 2 --warmup 1 > gpurun_out/r02c_ncu.log 2>&1; tail -2 gpurun_out/r02c_ncu.log
