"""The table shared by tests/golden/make_device_math_golden.py (reference side: oracle/_ref/libref_device_math.so, the reference CPU
backend's device_math.h compiled) and tests/test_device_math_reference.py (device side: the same Funcs through the IR -> CUDA
lowering on B200).  Names are `ir::Func` variants; the reference function each maps to is cpp.rs:520-640's table."""
import numpy as np

N = 256  # float4 / uint4 records per array

F4_UNARY = ["Abs", "Acos", "Acosh", "Asin", "Asinh", "Atan", "Atanh", "Cos", "Cosh", "Sin", "Sinh", "Tan", "Tanh", "Exp", "Exp2", "Exp10",
            "Log", "Log2", "Log10", "Sqrt", "Rsqrt", "Ceil", "Floor", "Fract", "Trunc", "Round", "Saturate", "Normalize", "Neg"]
F4_BINARY = ["Atan2", "Powf", "Copysign", "Min", "Max", "Step", "Add", "Sub", "Mul", "Div", "Rem"]
F4_TERNARY = ["Fma", "Clamp", "Lerp", "SmoothStep"]
F3_GEOMETRY = ["Cross", "Dot", "Length", "LengthSquared", "Distance", "Normalize", "Faceforward", "Reflect", "ReduceSum", "ReduceProd", "ReduceMin", "ReduceMax"]
MAT3 = ["Transpose", "Inverse", "MatMul", "MatCompMul", "OuterProduct", "Determinant", "MatVec"]
U4 = ["PopCount", "Clz", "Ctz", "Reverse", "Min", "Max", "Add", "Sub", "Mul", "Div", "Rem", "BitAnd", "BitOr", "BitXor", "BitNot", "Shl", "Shr"]
I4 = ["Abs", "Neg", "Min", "Max", "Div", "Rem", "Shr"]
F4_TO_U4 = ["IsNan", "IsInf", "CastU32", "CastI32", "Bitcast"]

# results that are a single correctly rounded IEEE operation (or pure selection / integer work) on both sides: bit-identical.
# Everything else is a libm (glibc vs libdevice) or multi-operation result: compared to a tolerance stated in the test.
EXACT_F = {"Abs", "Ceil", "Floor", "Trunc", "Round", "Saturate", "Neg", "Copysign", "Min", "Max", "Step", "Add", "Sub", "Mul", "Div", "Fma", "Clamp", "Sqrt",
           "Fract", "Rem", "Transpose", "MatCompMul", "OuterProduct", "ReduceMin", "ReduceMax"}


def inputs():
    """Seeded operands with the special values in front; ranges chosen so that no case is undefined behaviour on the C++ side
    (integer division by zero, INT_MIN / -1, shifts >= 32, float -> int casts out of range)."""
    rng = np.random.default_rng(0xD3714)
    special = np.array([0.0, -0.0, 1.0, -1.0, 0.5, -0.5, 2.0, -2.0, np.inf, -np.inf, np.nan, 1e-3, -1e-3, 3.5, -3.5, 0.25], np.float32)
    def f4(lo, hi):
        a = rng.uniform(lo, hi, (N, 4)).astype(np.float32)
        a.reshape(-1)[:special.size] = rng.permutation(special)
        return a
    d = {"fa": f4(-2.0, 2.0), "fb": f4(-2.0, 2.0), "fc": f4(-2.0, 2.0)}
    # well-conditioned 3x3 columns for Inverse / Determinant: identity-dominant
    ga, gb, gc = (rng.uniform(-1.0, 1.0, (N, 4)).astype(np.float32) for _ in range(3))
    ga[:, 0] += 3.0; gb[:, 1] += 3.0; gc[:, 2] += 3.0
    d.update(ga=ga, gb=gb, gc=gc)
    ua = rng.integers(0, 2 ** 32, (N, 4), dtype=np.uint64).astype(np.uint32)
    ua.reshape(-1)[:6] = [0, 1, 0x80000000, 0xFFFFFFFF, 0x00010000, 0x12345678]
    ub = rng.integers(1, 2 ** 32, (N, 4), dtype=np.uint64).astype(np.uint32)
    d.update(ua=ua, ub=ub, us=(ub % np.uint32(32)).astype(np.uint32))
    ia = rng.integers(-2 ** 31 + 1, 2 ** 31, (N, 4), dtype=np.int64).astype(np.int32)
    ib = rng.integers(2, 2 ** 20, (N, 4), dtype=np.int64).astype(np.int32) * rng.choice(np.array([-1, 1], np.int32), (N, 4))
    d.update(ia=ia, ib=ib, is_=(np.abs(ib) % 32).astype(np.int32))
    d["fcast"] = (rng.uniform(-1.0, 1.0, (N, 4)) * 1.0e6).astype(np.float32)
    return d
