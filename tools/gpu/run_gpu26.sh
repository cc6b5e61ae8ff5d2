#!/bin/bash
# session 26: PLOC builder — LBVH vs PLOC (radius sweep) on the C3 soup and on a terrain, then the GPU suite with PLOC as default
mkdir -p gpurun_out
LC_B200_BUILDER=lbvh timeout 300 python tools/trace_bench.py --tag lbvh --check 20000 2>&1 | tail -1
for r in 4 8 16; do LC_B200_BUILDER=ploc LC_B200_PLOC_RADIUS=$r timeout 300 python tools/trace_bench.py --tag ploc_r$r --check 20000 2>&1 | tail -2; done
LC_B200_BUILDER=lbvh timeout 300 python tools/trace_bench.py --tag lbvh --scene terrain --tris 5000000 2>&1 | tail -1
LC_B200_BUILDER=ploc timeout 300 python tools/trace_bench.py --tag ploc_r8 --scene terrain --tris 5000000 2>&1 | tail -2
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
