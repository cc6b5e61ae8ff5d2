#!/bin/bash
# usage: variant.sh <name> <extra nvcc flags...>  -> ../lib/variants/liblc_b200_<name>.so  (development aid for kernel sweeps)
set -e
cd "$(dirname "$0")"
name=$1; shift
out=../lib/variants; mkdir -p $out/$name
for f in device bvh_build radix_sort trace path_tracer; do
  extra=""; [ $f = path_tracer ] && extra="-fmad=false"
  /usr/local/cuda/bin/nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC,-fvisibility=hidden,-ffp-contract=off $extra "$@" -c $f.cu -o $out/$name/$f.o &
done
wait
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $out/liblc_b200_$name.so $out/$name/*.o -cudart static -lpthread
rm -rf $out/$name
echo built $out/liblc_b200_$name.so
