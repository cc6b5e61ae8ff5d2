#!/bin/bash
# round 2, session f: GPU suite after the curve refinement + lowering default change; curve bench; e2e chunk schedule sweep
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q --timeout 300 2>&1 | tail -12 | tee gpurun_out/r02f_tests.txt
for b in 0 1 3; do timeout 300 python tools/curve_bench.py --basis $b 2>/dev/null | tail -1; done | tee gpurun_out/r02f_curve_bench.jsonl
for chunk in 524288 1048576 2097152 4194304; do
  timeout 300 python bench.py --steps 10 --no-cpu --c5-spp 0 --e2e-chunk $chunk 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('chunk', $chunk, 'value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'e2e_batch', round(d['e2e_batch_entry']['value'],1), 'c2', round(d['dsl_path_tracer']['mrays_per_s'],1))"
done | tee gpurun_out/r02f_e2e_chunks.txt
