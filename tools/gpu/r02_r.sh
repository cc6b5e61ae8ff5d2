#!/bin/bash
# round 2, session r: bench with the two-thread e2e host program (uploads | everything else) and the new build
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 900 python bench.py --c5-spp 0 > gpurun_out/r02r_bench.json 2> gpurun_out/r02r_bench.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02r_bench.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','e2e','e2e_batch_entry','batch_entry','build','roofline','host_link','any_hit','gpu_launches'):
    print(k, json.dumps(d.get(k))[:400])
PY
tail -3 gpurun_out/r02r_bench.err
timeout 200 python tools/e2e_probe.py 2>&1 | tail -12 > gpurun_out/r02r_e2e_probe.txt
