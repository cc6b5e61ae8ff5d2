#!/bin/bash
# round 2, session y: C5 at N GPUs with more lanes per rank (dispatch tails of one lane filled by the others)
mkdir -p gpurun_out
N=${1:-8}; shift
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tools/c5_path_trace.py --spp 1024 "$@" 2>gpurun_out/r02y_c5_n$N.err | grep '^{' | tee -a gpurun_out/r02y_c5_n$N.jsonl | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print({k:d.get(k) for k in ('n_gpus','frame_ms','streams','tiles_per_rank','image_sha256')}); print([h['imbalance'] for h in d.get('balance_passes',[])])"
tail -2 gpurun_out/r02y_c5_n$N.err | cut -c1-200
