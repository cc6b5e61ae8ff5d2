#!/bin/bash
# round 2: bench.py under torchrun at N = 2, as the driver launches it
mkdir -p gpurun_out
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r02v_bench_n8.json 2> gpurun_out/r02v_bench_n8.err; echo "rc=$?"
python - <<'PY'
import json
try:
    d=json.loads([l for l in open('gpurun_out/r02v_bench_n8.json') if l.startswith('{')][-1])
    for k in ('value','n_gpus','ms_per_step','e2e','e2e_batch_entry','host_link'):
        print(k, json.dumps(d.get(k))[:420])
    c=d.get('c5_path_trace') or {}
    print('c5', {k:c.get(k) for k in ('frame_ms','spp','image_sha256','tiles_per_rank','imbalance_max_over_mean')})
except Exception as e:
    print("no bench line", e)
PY
tail -3 gpurun_out/r02v_bench_n8.err | cut -c1-300
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29552 bench.py --impl reference --gpus 8 --steps 2 --warmup 1 2>/dev/null | cut -c1-300
