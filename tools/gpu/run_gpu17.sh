#!/bin/bash
for c in 262144 524288 1048576 2097152; do LC_B200_HOST_CHUNK=$c timeout 300 python tools/e2e_bench.py 2>&1 | tail -1; done
timeout 900 python bench.py > gpurun_out/bench_c.json 2> gpurun_out/bench_c.err; echo "bench rc=$?"; cat gpurun_out/bench_c.json; tail -3 gpurun_out/bench_c.err
