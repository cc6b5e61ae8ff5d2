#!/bin/bash
# config C5 stand-alone at N GPUs (cost pass + balancing passes + one frame), knobs of tools/c5_path_trace.py from the command line:
#   gpurun --gpus 8 -- 'bash tools/gpu/c5.sh 8 --spp 1024 --recuts 0 --also-recuts 3'      (A/B of the in-frame re-cuts in one process)
#   gpurun -- 'bash tools/gpu/c5.sh 1 --spp 256 --streams 4'                                (lanes; the image SHA must not depend on them)
mkdir -p gpurun_out
N=${1:-2}; shift
if [ "$N" = 1 ]; then CMD="python tools/c5_path_trace.py"; else CMD="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tools/c5_path_trace.py"; fi
timeout 300 $CMD "$@" 2>gpurun_out/c5_n$N.err | grep '^{' | tee -a gpurun_out/c5_n$N.jsonl | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print({k:d.get(k) for k in ('config','n_gpus','spp','frame_ms','streams','recuts','tiles_per_rank','spp_per_pixel_ok','image_sha256')}); print([h['imbalance'] for h in d.get('balance_passes',[])], d.get('recuts_in_frame'))"
tail -2 gpurun_out/c5_n$N.err | cut -c1-200
