#!/bin/bash
# traversal work on one B200: the DSL call path against the batch kernel (tools/dsl_bench.py), kernel variants (csrc/variant.sh NAME flags -> LIBS="NAME ..."),
# the e2e timeline, and one `ncu --set full` capture (KERNEL=regex, default the kernel bench.py times) summarised by tools/ncu_summary.py
mkdir -p gpurun_out
T=${TAG:-trace}
timeout 600 python tools/dsl_bench.py c3 c2 2>/dev/null | tee gpurun_out/${T}_dsl.jsonl
for v in default $LIBS; do
  if [ $v = default ]; then unset LC_B200_LIB; else export LC_B200_LIB=luisa-compute-rs_b200/lib/variants/liblc_b200_$v.so; fi
  timeout 300 python tools/trace_bench.py --tag $v 2>&1 | tail -1
done | tee gpurun_out/${T}_variants.txt
unset LC_B200_LIB
timeout 200 python tools/e2e_probe.py 2>&1 | tail -12 | tee gpurun_out/${T}_e2e_probe.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:${KERNEL:-lc_kernel} -s 1 -c 1 -o gpurun_out/${T}_ncu -f python bench.py --profile --steps 2 --warmup 1 > gpurun_out/${T}_ncu.log 2>&1; tail -2 gpurun_out/${T}_ncu.log
python tools/ncu_summary.py gpurun_out/${T}_ncu.ncu-rep "${KERNEL:-lc_kernel}" > gpurun_out/${T}_ncu_full_summary.csv 2>/dev/null
