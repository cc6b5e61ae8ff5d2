#!/bin/bash
# bench.py under torchrun at N GPUs, as the driver launches it, followed by the reference arm:  gpurun --gpus N -- 'bash tools/gpu/bench_n.sh N'
mkdir -p gpurun_out
N=${1:-2}
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "rc=$?"
python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/bench_n$N.json') if l.startswith('{')][-1])
    for k in ('value','n_gpus','ms_per_step','e2e','e2e_batch_entry','host_link'):
        print(k, json.dumps(d.get(k))[:420])
    c=d.get('c5_path_trace') or {}
    print('c5', {k:c.get(k) for k in ('frame_ms','spp','image_sha256','tiles_per_rank','imbalance_max_over_mean')})
except Exception as e:
    print("no bench line", e)
PY
tail -3 gpurun_out/bench_n$N.err | cut -c1-300
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --impl reference --gpus $N --steps 2 --warmup 1 2>/dev/null | cut -c1-300
