#!/bin/bash
# C5 one-eighth share on one GPU (--emulate 0/8): samples per dispatch x streams — does a shorter thread shorten the tail?
mkdir -p gpurun_out
{
for spd in 64 32 16 8; do
  for st in 2 4; do
    timeout 300 python tools/c5_path_trace.py --emulate 0/8 --spp 64 --spp-per-dispatch $spd --streams $st 2>/dev/null | tail -1
  done
done
timeout 300 python tools/c5_path_trace.py --emulate 0/1 --spp 64 --spp-per-dispatch 32 --streams 2 2>/dev/null | tail -1
} | tee gpurun_out/r01z_c5_tail.txt
