#!/bin/bash
# session 28 (8 GPUs): scaling sanity — bench.py at N = 4, 8 and the C5 tile-sharded path tracer at N = 8
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for n in 4 8; do
  timeout 600 $TR --nproc-per-node $n --master-port $((29540 + n)) bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/r01n_bench_n$n.json 2> gpurun_out/bench_n$n.err; echo "bench n$n rc=$?"; cut -c1-330 gpurun_out/r01n_bench_n$n.json; tail -2 gpurun_out/bench_n$n.err
done
timeout 600 $TR --nproc-per-node 8 --master-port 29551 tools/c5_path_trace.py --spp 64 > gpurun_out/r01n_c5_pt_n8.json 2> gpurun_out/c5_n8.err; echo "c5 n8 rc=$?"; cat gpurun_out/r01n_c5_pt_n8.json; tail -2 gpurun_out/c5_n8.err
timeout 600 $TR --nproc-per-node 4 --master-port 29552 tools/c5_path_trace.py --spp 64 > gpurun_out/r01n_c5_pt_n4.json 2> gpurun_out/c5_n4.err; echo "c5 n4 rc=$?"; cat gpurun_out/r01n_c5_pt_n4.json; tail -2 gpurun_out/c5_n4.err
