#!/bin/bash
# round 2, session e (2 GPUs): bench.py under torchrun — weak-scaling C3 line + the C5 leg (cost-balanced partition, NCCL gather inside the timed region)
mkdir -p gpurun_out
N=${1:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02e_bench_n$N.json 2> gpurun_out/r02e_bench_n$N.err
echo "rc=$?"; tail -5 gpurun_out/r02e_bench_n$N.err; python - <<PY
import json
d=json.load(open("gpurun_out/r02e_bench_n$N.json"))
print({k:d[k] for k in ("value","n_gpus","ms_per_step")}, d.get("e2e",{}).get("value"), d.get("e2e_batch_entry",{}).get("value"))
c=d.get("c5_path_trace"); print({a:b for a,b in (c or {}).items() if a not in ("workload",)})
PY
