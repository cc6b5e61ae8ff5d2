#!/bin/bash
# compute-sanitizer over the rewritten build kernels (collapse work queue, 8-lane refit, optimistic climb, templated sort)
mkdir -p gpurun_out
SEL="test_c1_triangle or test_duplicate_centroids or test_rebuild_after or test_prefer_update or test_instances_transforms or test_empty_inputs or test_degenerate or test_hits_do_not_depend"
( echo "== memcheck =="; timeout 420 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_parity_gpu.py tests/test_curves.py -m gpu -x -q -k "$SEL or curve" --timeout 400 2>&1 | grep -v "^$" | tail -6; echo "rc=$?"
  echo "== racecheck (shared-memory hazards) =="; timeout 420 compute-sanitizer --tool racecheck --error-exitcode 3 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "test_c1_triangle or test_duplicate_centroids or test_prefer_update or test_instances_transforms" --timeout 400 2>&1 | grep -v "^$" | tail -6; echo "rc=$?"
  echo "== racecheck, 200 k-triangle soup build =="; timeout 300 compute-sanitizer --tool racecheck --error-exitcode 3 python tools/micro/build_once.py -200000 2 2>&1 | tail -4; echo "rc=$?"
) | tee gpurun_out/sanitizer.txt
