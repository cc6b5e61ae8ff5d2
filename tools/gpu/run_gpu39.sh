#!/bin/bash
timeout 1200 python -m pytest tests/test_parity_gpu.py -m gpu -x -q 2>&1 | grep -v "^  File\|^Extension" | tail -4
LC_B200_BUILDER=lbvh timeout 300 python tools/trace_bench.py --tag redux --check 20000 2>&1 | tail -1
LC_B200_BUILDER=lbvh timeout 300 python tools/trace_bench.py --tag redux --scene terrain --tris 20000000 --rays 8388608 2>&1 | tail -1
timeout 300 python tools/trace_bench.py --tag redux_auto --scene terrain --tris 5000000 2>&1 | tail -1
