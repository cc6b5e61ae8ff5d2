#!/bin/bash
# session 27: GPU suite with the builder tests; auto choice on soup / terrain; C4-size build times per builder
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -s 2>&1 | grep -v "^  File\|^Extension" | tail -6
( timeout 300 python tools/trace_bench.py --tag auto 2>&1 | tail -1
  timeout 300 python tools/trace_bench.py --tag auto --scene terrain --tris 5000000 2>&1 | tail -1
  LC_B200_BUILDER=lbvh timeout 300 python tools/trace_bench.py --tag lbvh --scene terrain --tris 20000000 --rays 8388608 2>&1 | tail -1
  LC_B200_BUILDER=ploc timeout 300 python tools/trace_bench.py --tag ploc --scene terrain --tris 20000000 --rays 8388608 2>&1 | tail -1 ) | tee gpurun_out/r01m_builder_sweep.txt
