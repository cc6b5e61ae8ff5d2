#!/bin/bash
mkdir -p gpurun_out
for cfg in "16 1" "8 1" "8 2" "16 2" "8 4"; do set -- $cfg
  timeout 600 python tools/c5_path_trace.py --spp 64 --block $1 --streams $2 2> gpurun_out/c5_var.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('block',d['block'],'streams',d['streams'],'render_ms',d['render_ms'],'mrays',d['mrays_per_s'],d['image_sha256'][:12])"; tail -1 gpurun_out/c5_var.err
done
