#!/bin/bash
# path regeneration vs nested loops on the C5 kernel, one GPU
mkdir -p gpurun_out
for mode in "--regenerate" ""; do
  for bl in 8 16; do
    echo "== mode=[$mode] block=$bl" >> gpurun_out/c5_regen.log
    timeout 600 python tools/c5_path_trace.py --spp 64 --block $bl $mode 2>/dev/null | tail -1 >> gpurun_out/c5_regen.log
  done
done
cat gpurun_out/c5_regen.log
