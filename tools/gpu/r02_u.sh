#!/bin/bash
# round 2, session u: C5 at N GPUs: cost map from per-tile ray counts, frame without re-cuts, then the same frame with 3 re-cuts (A/B in one process)
mkdir -p gpurun_out
N=${1:-2}; SPP=${2:-256}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tools/c5_path_trace.py --spp $SPP --recuts 0 --also-recuts 3 2>gpurun_out/r02u_c5_n$N.err | grep '^{' | tee -a gpurun_out/r02u_c5_n$N.jsonl | cut -c1-1900
tail -2 gpurun_out/r02u_c5_n$N.err | cut -c1-300
