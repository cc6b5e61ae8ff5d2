#!/bin/bash
# round 2, session x: compute-sanitizer over the rewritten build kernels (collapse work queue, 8-lane refit, optimistic climb, templated sort) + build timing after the k_pack_tris change
mkdir -p gpurun_out
SEL="test_c1_triangle or test_duplicate_centroids or test_rebuild_after or test_prefer_update or test_instances_transforms or test_empty_inputs or test_degenerate or test_hits_do_not_depend"
( echo "== memcheck =="; timeout 420 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_parity_gpu.py tests/test_curves.py -m gpu -x -q -k "$SEL or curve" --timeout 400 2>&1 | grep -v "^$" | tail -6; echo "rc=$?"
  echo "== racecheck (shared-memory hazards) =="; timeout 420 compute-sanitizer --tool racecheck --error-exitcode 3 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "test_c1_triangle or test_duplicate_centroids or test_prefer_update or test_instances_transforms" --timeout 400 2>&1 | grep -v "^$" | tail -6; echo "rc=$?"
  echo "== racecheck, 200 k-triangle soup build =="; timeout 300 compute-sanitizer --tool racecheck --error-exitcode 3 python tools/micro/build_once.py -200000 2 2>&1 | tail -4; echo "rc=$?"
) | tee gpurun_out/r02x_sanitizer.txt
timeout 200 python tools/trace_bench.py --tag x --rays 4194304 --check 65536 2>&1 | tail -1 | tee gpurun_out/r02x_build.txt
LC_B200_BUILDER=lbvh timeout 200 python tools/trace_bench.py --tag x-terrain20M-lbvh --scene terrain --tris 20000000 --rays 8388608 2>&1 | tail -1 | tee -a gpurun_out/r02x_build.txt
timeout 200 python tools/micro/rebuild_probe.py 2>&1 | sed -n 2,6p | tee -a gpurun_out/r02x_build.txt
