#!/bin/bash
# round 2, session a: wavefront lowering — parity on the GPU, then throughput of the DSL call path against the batch kernel, knobs swept
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_wavefront_lowering.py tests/test_ir_lowering.py -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/r02a_tests.txt
timeout 600 python tools/dsl_bench.py c3 > gpurun_out/r02a_dsl_c3.jsonl 2> gpurun_out/r02a_dsl_c3.err; tail -3 gpurun_out/r02a_dsl_c3.err; cat gpurun_out/r02a_dsl_c3.jsonl
timeout 600 python tools/dsl_bench.py c2 > gpurun_out/r02a_dsl_c2.jsonl 2> gpurun_out/r02a_dsl_c2.err; tail -3 gpurun_out/r02a_dsl_c2.err; cat gpurun_out/r02a_dsl_c2.jsonl
for v in nostream mb6 mb5; do LC_B200_LIB=luisa-compute-rs_b200/lib/variants/liblc_b200_$v.so timeout 300 python tools/trace_bench.py --tag $v 2>&1 | tail -1; done | tee gpurun_out/r02a_variants.txt
timeout 300 python tools/trace_bench.py --tag default 2>&1 | tail -1 | tee -a gpurun_out/r02a_variants.txt
