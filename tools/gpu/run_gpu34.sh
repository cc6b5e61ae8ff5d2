#!/bin/bash
( for p in 6 5 4 3; do LC_B200_SORT_PASSES=$p timeout 300 python tools/trace_bench.py --tag passes$p 2>&1 | tail -1; done
  for p in 6 5 4; do LC_B200_SORT_PASSES=$p timeout 300 python tools/trace_bench.py --tag passes$p --scene terrain --tris 20000000 --rays 8388608 2>&1 | tail -1; done
  timeout 300 python tools/trace_bench.py --tag auto_passes 2>&1 | tail -1 ) | tee gpurun_out/r01r_sort_passes.txt
