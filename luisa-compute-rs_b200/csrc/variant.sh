#!/bin/bash
# usage: variant.sh <name> <extra nvcc flags...>  -> ../lib/variants/liblc_b200_<name>.so  (development aid for kernel sweeps)
set -e
cd "$(dirname "$0")"
name=$1; shift
out=../lib/variants; mkdir -p $out/$name
make -s ../lib/embedded_headers.inc
for f in device bvh_build radix_sort trace path_tracer shader ir_lower; do
  extra=""; [ $f = path_tracer ] && extra="-fmad=false"
  src=$f.cu; [ $f = ir_lower ] && src="-x cu ir_lower.cpp"
  /usr/local/cuda/bin/nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -I../lib -Xcompiler -fPIC,-fvisibility=hidden,-ffp-contract=off -Xptxas -v $extra "$@" -c $src -o $out/$name/$f.o 2> $out/$name/$f.log &
done
wait
grep -A1 "k_traceILi0ELb0" $out/$name/trace.log | grep -i "spill\|registers" | head -4
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $out/liblc_b200_$name.so $out/$name/*.o -cudart static -lpthread -ldl
rm -rf $out/$name
echo built $out/liblc_b200_$name.so
