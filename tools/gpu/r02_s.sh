#!/bin/bash
# round 2, session s: bench with the two-thread e2e host program (waits follow signals), hard timeouts
mkdir -p gpurun_out
timeout 420 python bench.py --c5-spp 0 > gpurun_out/r02s_bench.json 2> gpurun_out/r02s_bench.err; echo "bench rc=$?"; python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r02s_bench.json').read().strip().splitlines()[-1])
    for k in ('value','ms_per_step','e2e','e2e_batch_entry','batch_entry','build','roofline','host_link','any_hit','gpu_launches','cpu_baseline'):
        print(k, json.dumps(d.get(k))[:420])
except Exception as e:
    print("no bench line", e)
PY
tail -3 gpurun_out/r02s_bench.err
