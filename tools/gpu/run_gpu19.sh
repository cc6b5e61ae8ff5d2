#!/bin/bash
V=luisa-compute-rs_b200/lib/variants
for v in leaf2 leaf1; do LC_B200_LIB=$V/liblc_b200_$v.so timeout 300 python tools/trace_bench.py --tag $v --check 20000 2>&1 | tail -1; done
timeout 300 python tools/trace_bench.py --tag leaf3 2>&1 | tail -1
