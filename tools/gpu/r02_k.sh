#!/bin/bash
# C5 stand-alone at N GPUs (probe + balancing passes + one frame), knobs from the command line
mkdir -p gpurun_out
N=${1:-2}; shift
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tools/c5_path_trace.py "$@" 2>gpurun_out/r02k_c5_n$N.err | tail -1 | tee -a gpurun_out/r02k_c5_n$N.jsonl
tail -3 gpurun_out/r02k_c5_n$N.err
