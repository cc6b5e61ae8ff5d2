#!/bin/bash
# round 2, session n: warp-autonomous collapse (no CTA barriers; termination = every primitive placed), CTA size / occupancy variants, sort tile variants
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_curves.py -m gpu -x -q --timeout 300 2>&1 | tail -4 | tee gpurun_out/r02n_tests.txt
for v in default c256 c128x10 c128x12 s8 s16; do
  if [ $v = default ]; then unset LC_B200_LIB; else export LC_B200_LIB=luisa-compute-rs_b200/lib/variants/liblc_b200_$v.so; fi
  timeout 300 python tools/trace_bench.py --tag $v --rays 4194304 --check 65536 2>&1 | tail -1
done | tee gpurun_out/r02n_variants.txt
unset LC_B200_LIB
LC_B200_BUILDER=lbvh timeout 300 python tools/trace_bench.py --tag n-terrain20M-lbvh --scene terrain --tris 20000000 --rays 8388608 2>&1 | tail -1 | tee -a gpurun_out/r02n_variants.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02n_build_1M_launches.csv python tools/micro/build_once.py -1000000 3 > gpurun_out/r02n_ncu1.log 2>&1
python tools/launch_summary.py gpurun_out/r02n_build_1M_launches.csv | tee gpurun_out/r02n_build_1M_launches_summary.csv
ncu --set full --clock-control none --import-source on -k regex:'k_collapse' -s 1 -c 1 -o gpurun_out/r02n_collapse -f python tools/micro/build_once.py -1000000 2 > gpurun_out/r02n_ncu3.log 2>&1; tail -1 gpurun_out/r02n_ncu3.log
