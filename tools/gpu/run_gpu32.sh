#!/bin/bash
V=luisa-compute-rs_b200/lib/variants
timeout 300 python tools/trace_bench.py --tag base 2>&1 | tail -1
for v in mb6 ss8 ss12 mb6ss12 mb6ss24; do LC_B200_LIB=$V/liblc_b200_$v.so timeout 300 python tools/trace_bench.py --tag $v 2>&1 | tail -1; done
