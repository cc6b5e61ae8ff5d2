#!/bin/bash
# session 24 (2 GPUs): C5 tile-sharded path tracing — N-independence of the image at a reduced size, then 4K throughput at N = 1 and 2;
# default bench at N = 2
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 python tools/c5_path_trace.py --width 960 --height 540 --spp 32 --nx 400 > gpurun_out/c5_small_n1.json 2> gpurun_out/c5_small_n1.err; echo "small n1 rc=$?"; cat gpurun_out/c5_small_n1.json; tail -3 gpurun_out/c5_small_n1.err
timeout 600 $TR --nproc-per-node 2 --master-port 29531 tools/c5_path_trace.py --width 960 --height 540 --spp 32 --nx 400 > gpurun_out/c5_small_n2.json 2> gpurun_out/c5_small_n2.err; echo "small n2 rc=$?"; cat gpurun_out/c5_small_n2.json; tail -3 gpurun_out/c5_small_n2.err
timeout 900 python tools/c5_path_trace.py --spp 64 > gpurun_out/r01k_c5_pt_n1.json 2> gpurun_out/c5_n1.err; echo "4k n1 rc=$?"; cat gpurun_out/r01k_c5_pt_n1.json; tail -3 gpurun_out/c5_n1.err
timeout 900 $TR --nproc-per-node 2 --master-port 29532 tools/c5_path_trace.py --spp 64 > gpurun_out/r01k_c5_pt_n2.json 2> gpurun_out/c5_n2.err; echo "4k n2 rc=$?"; cat gpurun_out/r01k_c5_pt_n2.json; tail -3 gpurun_out/c5_n2.err
timeout 900 $TR --nproc-per-node 2 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r01k_bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench n2 rc=$?"; cat gpurun_out/r01k_bench_n2.json; tail -3 gpurun_out/bench_n2.err
