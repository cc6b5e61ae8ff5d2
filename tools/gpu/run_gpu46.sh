#!/bin/bash
# host pipeline: CTAs per chunk launch (LC_B200_HOST_GRID) x chunk schedule
mkdir -p gpurun_out
{
for g in 0 777 592 518 444; do
  for mx in 1048576 2097152; do
    echo -n "[grid $g max $mx] "; LC_B200_HOST_GRID=$g LC_B200_HOST_CHUNK_MAX=$mx timeout 200 python tools/e2e_bench.py 2>/dev/null | tail -1
  done
done
} | tee gpurun_out/r01z_e2e_grid.txt
