#!/bin/bash
for c in 1 4 16 64 255; do
  timeout 300 python tools/c5_path_trace.py --spp 64 --block 8 --streams 4 --spp-per-dispatch 16 --chunk $c --emulate 0/8 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('chunk $c tiles',d['tiles'],'ms',d['render_ms'],'Mrays/s',round(d['rays']/d['render_ms']/1e3,1))"
done
