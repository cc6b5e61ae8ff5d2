#!/bin/bash
V=luisa-compute-rs_b200/lib/variants
export LC_B200_BUILDER=lbvh
for v in base cmb5 cmb6 cmb8; do
  lib=""; [ $v != base ] && lib=$V/liblc_b200_$v.so
  LC_B200_LIB=$lib timeout 300 python tools/micro/build_once.py -1000000 6 2>&1 | tail -1 | sed "s/^/[$v 1M] /"
  LC_B200_LIB=$lib timeout 300 python tools/micro/build_once.py 3164 4 2>&1 | tail -1 | sed "s/^/[$v 20M] /"
done
