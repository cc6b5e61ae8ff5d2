export LC_B200_DEBUG_ACCEL=1
timeout 300 python -m pytest tests/test_curves.py tests/test_ir_lowering.py -m gpu -q -k "curve or test_procedural_primitives_through_ray_query or test_ray_query_ir_kernel" 2>&1 | grep -a "DBG\|passed\|failed" | tail -8
echo ==== passing
timeout 300 python -m pytest tests/test_curves.py tests/test_ir_lowering.py -m gpu -q -k "test_procedural_primitives_through_ray_query or test_ray_query_ir_kernel" 2>&1 | grep -a "DBG\|passed\|failed" | tail -5
