#!/bin/bash
# round 2, session g: 96-byte nodes (8-bit planes, the new default) against the 128-byte layout (variant nb16): parity suite on the default build, then
# Mrays/s + nodes per ray on soup and terrain
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q --timeout 300 2>&1 | tail -12 | tee gpurun_out/r02g_tests.txt
for v in default nb16; do
  if [ $v = default ]; then unset LC_B200_LIB; else export LC_B200_LIB=luisa-compute-rs_b200/lib/variants/liblc_b200_$v.so; fi
  timeout 300 python tools/trace_bench.py --tag $v 2>&1 | tail -1
  timeout 300 python tools/trace_bench.py --tag $v --scene terrain --tris 5000000 2>&1 | tail -1
done | tee gpurun_out/r02g_node_bits.txt
unset LC_B200_LIB
timeout 600 python bench.py --steps 10 --no-cpu --c5-spp 64 2>/dev/null > gpurun_out/r02g_bench.json; python - <<'PY'
import json
d=json.load(open("gpurun_out/r02g_bench.json"))
print("value", d["value"], "batch", d["batch_entry"]["value"], "nodes/ray", d["roofline"]["nodes_per_ray"], "tris/ray", d["roofline"]["tris_per_ray"], "frac", d["roofline"]["frac"], "bvh_bytes", d["build"]["bvh_bytes"], "blas_ms", d["build"]["blas_ms"], "any", d["any_hit"], "c2", d["dsl_path_tracer"]["mrays_per_s"], "c5 ms", d["c5_path_trace"]["frame_ms"], "e2e", d["e2e"]["value"])
PY
