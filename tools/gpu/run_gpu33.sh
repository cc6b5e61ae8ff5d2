#!/bin/bash
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "builder or full_size_properties_c4 or refit or duplicate" 2>&1 | grep -v "^  File\|^Extension" | tail -4
LC_B200_BUILDER=ploc timeout 300 python tools/trace_bench.py --tag ploc --scene terrain --tris 20000000 --rays 8388608 2>&1 | tail -1
LC_B200_BUILDER=ploc timeout 300 python tools/trace_bench.py --tag ploc --scene terrain --tris 5000000 2>&1 | tail -1
LC_B200_BUILDER=ploc timeout 300 python tools/trace_bench.py --tag ploc 2>&1 | tail -1
