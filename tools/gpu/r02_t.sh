#!/bin/bash
# round 2, session t: C5 with in-frame re-cuts at N = 2 (256 spp): N=1 reference SHA, N=2 without and with re-cuts
mkdir -p gpurun_out
run() { N=$1; shift; timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tools/c5_path_trace.py "$@" 2>gpurun_out/r02t_c5.err | tail -1 | tee -a gpurun_out/r02t_c5.jsonl | cut -c1-1500; tail -2 gpurun_out/r02t_c5.err | cut -c1-300; }
run 1 --spp 256
run 2 --spp 256 --recuts 0
run 2 --spp 256 --recuts 6
