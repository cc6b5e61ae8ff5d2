#!/bin/bash
# C5 eighths: tiles per round-robin run along the Morton curve (--chunk) x rank: locality vs balance
mkdir -p gpurun_out
{
for ch in 1 4 16 64; do
  for r in 0 2 5 7; do
    echo -n "chunk $ch rank $r: "; timeout 300 python tools/c5_path_trace.py --emulate $r/8 --spp 64 --spp-per-dispatch 16 --streams 4 --chunk $ch 2>/dev/null | tail -1
  done
done
} | tee gpurun_out/r01z_c5_chunk.txt
