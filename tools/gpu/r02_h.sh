#!/bin/bash
# round 2, session h: occupancy / refill knobs again with 96-byte nodes (batch kernel and wavefront-lowered C3 kernel), then ncu --set full of k_trace
mkdir -p gpurun_out
for v in default mb7 mb5; do
  if [ $v = default ]; then unset LC_B200_LIB; else export LC_B200_LIB=luisa-compute-rs_b200/lib/variants/liblc_b200_$v.so; fi
  timeout 300 python tools/trace_bench.py --tag $v 2>&1 | tail -1
done | tee gpurun_out/r02h_occupancy.txt
export LC_B200_LIB=luisa-compute-rs_b200/lib/variants/liblc_b200_mb7.so
for f in 3 4 8 10; do LC_B200_TRACE_TUNE=$f timeout 300 python tools/trace_bench.py --tag "mb7 fetch_min=$f" 2>&1 | tail -1; done | tee -a gpurun_out/r02h_occupancy.txt
unset LC_B200_LIB
for f in 4 8 10; do LC_B200_TRACE_TUNE=$f timeout 300 python tools/trace_bench.py --tag "mb6 fetch_min=$f" 2>&1 | tail -1; done | tee -a gpurun_out/r02h_occupancy.txt
timeout 400 python tools/dsl_bench.py c3 --configs 0:5:4,0:5:6,0:5:8,0:5:12,0:6:6,0:6:8,0:4:8,0:7:8 2>/dev/null | tee gpurun_out/r02h_dsl_c3.jsonl
ncu --set full --clock-control none --import-source on -k regex:k_trace -s 2 -c 1 -o gpurun_out/r02h_k_trace -f python tools/trace_bench.py --reps 2 > gpurun_out/r02h_ncu.log 2>&1; tail -2 gpurun_out/r02h_ncu.log
