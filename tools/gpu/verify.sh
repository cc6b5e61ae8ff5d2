#!/bin/bash
# full verification on one B200 (TAG=name of the output files, default "verify"): GPU suite, smoke, bench (default flags: N = 1, C5 leg included), reference arm, ncu launch list of the bench command
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 2>&1 | tail -4 | tee gpurun_out/${TAG:-verify}_tests.txt
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/${TAG:-verify}_bench.json 2> gpurun_out/${TAG:-verify}_bench.err; echo "bench rc=$?"; python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/${TAG:-verify}_bench.json').read().strip().splitlines()[-1])
    for k in ('value','ms_per_step','e2e','e2e_batch_entry','batch_entry','build','roofline','any_hit','gpu_launches','cpu_baseline','parity_sample','clocks'):
        print(k, json.dumps(d.get(k))[:360])
    c=d.get('c5_path_trace') or {}
    print('c5', {k:c.get(k) for k in ('frame_ms','spp','image_sha256','mrays_per_s')})
    print('c2', json.dumps(d.get('dsl_path_tracer'))[:300])
except Exception as e:
    print("no bench line", e)
PY
tail -3 gpurun_out/${TAG:-verify}_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null > gpurun_out/${TAG:-verify}_bench_reference.json; cut -c1-400 gpurun_out/${TAG:-verify}_bench_reference.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG:-verify}_launches.csv python bench.py --profile --steps 2 --warmup 1 > gpurun_out/${TAG:-verify}_profile.log 2>&1
python tools/launch_summary.py gpurun_out/${TAG:-verify}_launches.csv | tee gpurun_out/${TAG:-verify}_launches_summary.csv
