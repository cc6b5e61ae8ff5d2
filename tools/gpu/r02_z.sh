#!/bin/bash
# round 2, session z: is the image independent of the number of lanes?  N = 1, 256 spp, streams 1 / 2 / 4 (expected SHA eabe67c8...)
mkdir -p gpurun_out
for s in 1 2 4; do
  timeout 200 python tools/c5_path_trace.py --spp 256 --streams $s 2>gpurun_out/r02z.err | grep '^{' | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print({k:d.get(k) for k in ('streams','frame_ms','spp_per_pixel_ok','image_sha256','rays')})"
done | tee gpurun_out/r02z_lanes.txt
tail -2 gpurun_out/r02z.err | cut -c1-200
