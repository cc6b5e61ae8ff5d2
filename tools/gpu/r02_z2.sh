#!/bin/bash
# session z2: lanes x ranks: N = 2, 256 spp, streams 4 with and without the balancing passes (expected SHA eabe67c8...)
mkdir -p gpurun_out
for extra in "--balance-passes 0" "--balance-passes 6"; do
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/c5_path_trace.py --spp 256 --streams 4 $extra 2>gpurun_out/r02z2.err | grep '^{' | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print({k:d.get(k) for k in ('streams','frame_ms','spp_per_pixel_ok','image_sha256','tiles_per_rank')}); print([ (h['imbalance'],h['ms_per_rank']) for h in d.get('balance_passes',[])])"
done | tee gpurun_out/r02z2_lanes.txt
tail -2 gpurun_out/r02z2.err | cut -c1-200
