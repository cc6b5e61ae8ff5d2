#!/bin/bash
timeout 900 python -m pytest tests/test_ir_lowering.py -m gpu -x -q 2>&1 | grep -v "^  File\|^Extension" | tail -3
for v in 1 0; do
  if [ $v = 1 ]; then export LC_B200_RT_LOCAL_STACK=1; else unset LC_B200_RT_LOCAL_STACK; fi
  echo "== local stack: $v"
  timeout 600 python tools/c2_bench.py 2>/dev/null | grep lowered_poly | cut -c1-200
  timeout 600 python tools/c5_path_trace.py --spp 64 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('c5 render_ms',d['render_ms'],'mrays',d['mrays_per_s'],d['image_sha256'][:12])"
done
