#!/bin/bash
# emulate rank 0 of 8 on one GPU: where does the per-rank slowdown at N = 8 come from?
for cfg in "8 2 32" "8 1 32" "8 2 64" "8 4 16" "8 2 8" "16 2 32" "4 2 32"; do set -- $cfg
  timeout 300 python tools/c5_path_trace.py --spp 64 --block $1 --streams $2 --spp-per-dispatch $3 --emulate 0/8 2>/dev/null
done
timeout 300 python tools/c5_path_trace.py --spp 64 --block 8 --streams 2 --emulate 0/1 2>/dev/null
