"""IR builtins against the reference CPU backend's own kernel library.

`tests/golden/device_math_reference.npz` holds what device_math.h — the header every JIT-ed kernel of the reference `cpu` device is
compiled with (cpu/codegen/cpp.rs:2064-2080) — returns for the operands of tests/device_math_cases.py; it was produced by compiling
that header where it lies (oracle/Makefile `ref_device_math`, tests/golden/make_device_math_golden.py).  The GPU tests run the same
`ir::Func`s through create_shader (IR -> CUDA lowering, compiled without fast-math: -fmad=false, IEEE div / sqrt, no FTZ) and compare:
bit-identical where the result is one correctly rounded operation, selection or integer work (`EXACT_F`, all integer tables); for libm
functions (glibc there, libdevice here) and multi-operation formulas a tolerance, stated per group below.  NaNs must appear in the
same places.  One documented divergence: `Func::Reverse` (see test_reverse_is_bit_reversal)."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import device_math_cases as cases  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "device_math_reference.npz")
REF_LIB = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "oracle", "_ref", "libref_device_math.so")


def test_golden_file_is_complete():
    g = np.load(GOLDEN)
    want = ["f1_" + n for n in cases.F4_UNARY] + ["f2_" + n for n in cases.F4_BINARY] + ["f3_" + n for n in cases.F4_TERNARY] + \
           ["g3_" + n for n in cases.F3_GEOMETRY] + ["m3_" + n for n in cases.MAT3] + ["u4_" + n for n in cases.U4] + \
           ["i4_" + n for n in cases.I4] + ["fu_" + n for n in cases.F4_TO_U4]
    assert sorted(g.files) == sorted(want)
    # spot values any reader can verify by hand
    d = cases.inputs()
    i = int(np.argwhere(d["ua"].reshape(-1) == 0x12345678)[0, 0])
    assert g["u4_PopCount"].reshape(-1)[i] == 13 and g["u4_Clz"].reshape(-1)[i] == 3 and g["u4_Ctz"].reshape(-1)[i] == 3
    assert g["u4_Reverse"].reshape(-1)[i] == 0x78563412   # the reference `cpu` device swaps bytes (cpu_prelude.h __brev)


@pytest.mark.skipif(not os.path.exists(REF_LIB), reason="oracle/_ref/libref_device_math.so is built only where the reference tree is present")
def test_committed_vectors_are_what_the_compiled_reference_returns():
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import make_device_math_golden as mk
    fresh, g = mk.compute(REF_LIB), np.load(GOLDEN)
    for k in g.files:
        assert np.array_equal(fresh[k].view(np.uint32), g[k].view(np.uint32)), k


def test_texture_sampling_restatement_equals_the_reference_vectors():
    """cpu_texture.h (lc_texture_2d_sample: point / bilinear x edge / repeat / mirror / zero, on the `cpu` device's 4 x 4-blocked image
    storage) compiled -> tests/golden/texture_sample_reference.npz; the numpy restatement the GPU sampling test compares with bit for
    bit (test_ir_lowering.np_sample2d) returns exactly those values, so device == restatement == compiled reference."""
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import make_device_math_golden as mk
    import test_ir_lowering as til
    g = np.load(os.path.join(os.path.dirname(GOLDEN), "texture_sample_reference.npz"))
    img, uv = mk.texture_inputs()
    for filt in (0, 1):
        for address in range(4):
            want = g["tex2d_%d_%d" % (filt, address)]
            got = til.np_sample2d(img, uv, filt, address).astype(np.float32)
            assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), (filt, address)
    if os.path.exists(REF_LIB):
        fresh = dict(mk.compute_texture(REF_LIB), **mk.compute_texture3d(REF_LIB))
        assert sorted(fresh) == sorted(g.files)
        for k in g.files:
            assert np.array_equal(fresh[k].view(np.uint32), g[k].view(np.uint32)), k


# ---- device side ----------------------------------------------------------------------------------------------------------------------
def _build(names, n_in, in_ty, out_ty, emit, rows_per_item=1):
    """One kernel per table: out[k * rows * N + rows * i + r] = F_k(a[i], b[i], c[i])."""
    from luisa_compute_rs_b200 import ir
    k = ir.KernelBuilder(block_size=(64, 1, 1))
    it, ot = getattr(k, in_ty), getattr(k, out_ty)
    ins = [k.arg_buffer(it) for _ in range(n_in)]
    out = k.arg_buffer(ot)

    def body():
        i = k.dispatch_id().x
        vals = [b.read(i) for b in ins]
        for idx, name in enumerate(names):
            res = emit(k, name, vals)
            res = res if isinstance(res, list) else [res]
            assert len(res) == rows_per_item
            for r, v in enumerate(res):
                out.write(i * rows_per_item + (idx * rows_per_item * cases.N + r), v)
    k.body(body)
    k.finish()
    return k


def _run(device, k, arrays, n_funcs, out_dtype, rows_per_item=1):
    bufs = [device.create_buffer_from_array(a) for a in arrays]
    out = device.create_buffer(n_funcs * rows_per_item * cases.N, 16, 16)
    sh = device.create_shader(C.addressof(k.km), keep=k)   # fast_math=False
    sh.dispatch((cases.N,), *bufs, out)
    got = out.view().to_numpy(out_dtype).reshape(n_funcs, rows_per_item * cases.N, 4).copy()
    for r in bufs + [out, sh]:
        r.destroy()
    return got


def _check_float(name, got, want, rtol, atol, lanes=4):
    got, want = got[:, :lanes], want[:, :lanes]
    assert np.array_equal(np.isnan(got), np.isnan(want)), "%s: NaNs in different places" % name
    if rtol == 0:
        same = (got.view(np.uint32) == want.view(np.uint32)) | (np.isnan(got) & np.isnan(want))
        assert same.all(), "%s: %d of %d results differ in bits, e.g. %r vs %r" % (name, (~same).sum(), same.size, got[~same][:3], want[~same][:3])
    else:
        ok = np.isclose(got.astype(np.float64), want.astype(np.float64), rtol=rtol, atol=atol, equal_nan=True)
        assert ok.all(), "%s: worst %r vs %r" % (name, got[~ok][:3], want[~ok][:3])


LIBM_RTOL = 4e-7      # ~3 ulp: libdevice's documented worst case for the functions below is 2 ulp, glibc's 1
FORMULA_RTOL = 2e-6   # multi-operation formulas (lerp, smoothstep, dot, cross, inverse ...): same formula, possibly another association


def _f(F, name):
    return getattr(F, name)


@pytest.mark.gpu
def test_float4_unary_builtins(device):
    from luisa_compute_rs_b200.ir import Func
    g, d = np.load(GOLDEN), cases.inputs()
    k = _build(cases.F4_UNARY, 1, "f324", "f324", lambda k, n, v: (-v[0]) if n == "Neg" else v[0].unary(_f(Func, n)))
    got = _run(device, k, [d["fa"]], len(cases.F4_UNARY), np.float32)
    for i, n in enumerate(cases.F4_UNARY):
        exact = n in cases.EXACT_F
        _check_float(n, got[i], g["f1_" + n], 0 if exact else (FORMULA_RTOL if n == "Normalize" else LIBM_RTOL), 0 if exact else 1e-7)


@pytest.mark.gpu
def test_float4_binary_and_ternary_builtins(device):
    from luisa_compute_rs_b200.ir import Func
    g, d = np.load(GOLDEN), cases.inputs()
    ops = {"Add": lambda a, b: a + b, "Sub": lambda a, b: a - b, "Mul": lambda a, b: a * b, "Div": lambda a, b: a / b, "Rem": lambda a, b: a % b}

    def emit2(k, n, v):
        return ops[n](v[0], v[1]) if n in ops else k.call(_f(Func, n), [v[0], v[1]], k.f324)
    got = _run(device, _build(cases.F4_BINARY, 2, "f324", "f324", emit2), [d["fa"], d["fb"]], len(cases.F4_BINARY), np.float32)
    for i, n in enumerate(cases.F4_BINARY):
        exact = n in cases.EXACT_F
        _check_float(n, got[i], g["f2_" + n], 0 if exact else LIBM_RTOL, 0 if exact else 1e-7)
    got = _run(device, _build(cases.F4_TERNARY, 3, "f324", "f324", lambda k, n, v: k.call(_f(Func, n), v, k.f324)), [d["fa"], d["fb"], d["fc"]], len(cases.F4_TERNARY), np.float32)
    for i, n in enumerate(cases.F4_TERNARY):
        exact = n in cases.EXACT_F
        _check_float(n, got[i], g["f3_" + n], 0 if exact else FORMULA_RTOL, 0 if exact else 1e-6)


@pytest.mark.gpu
def test_float3_geometry_and_matrix_builtins(device):
    from luisa_compute_rs_b200.ir import Func
    g, d = np.load(GOLDEN), cases.inputs()
    scalar = {"Dot", "Length", "LengthSquared", "Distance", "ReduceSum", "ReduceProd", "ReduceMin", "ReduceMax"}
    arity = {"Cross": 2, "Dot": 2, "Distance": 2, "Reflect": 2, "Faceforward": 3}

    def xyz(k, v):
        return k.vec(k.f323, v.x, v.y, v.z)

    def emit(k, n, v):
        a = [xyz(k, x) for x in v][:arity.get(n, 1)]
        r = k.call(_f(Func, n), a, k.f32 if n in scalar else k.f323)
        return k.vec(k.f324, r, r, r, k.f(0.0)) if n in scalar else k.vec(k.f324, r.x, r.y, r.z, k.f(0.0))
    got = _run(device, _build(cases.F3_GEOMETRY, 3, "f324", "f324", emit), [d["fa"], d["fb"], d["fc"]], len(cases.F3_GEOMETRY), np.float32)
    for i, n in enumerate(cases.F3_GEOMETRY):
        exact = n in cases.EXACT_F
        _check_float(n, got[i], g["g3_" + n], 0 if exact else FORMULA_RTOL, 0 if exact else 2e-6, lanes=3)

    def emit_m(k, n, v):
        c = [xyz(k, x) for x in v]
        m = k.call(Func.Mat3, c, k.matrix(3))
        mt = k.call(Func.Transpose, [m], k.matrix(3))
        if n == "Determinant":
            s = k.call(Func.Determinant, [m], k.f32)
            return [k.vec(k.f324, s, s, s, k.f(0.0))] * 3
        if n == "MatVec":
            r = m * c[2]
            return [k.vec(k.f324, r.x, r.y, r.z, k.f(0.0))] * 3
        r = {"Transpose": lambda: mt, "Inverse": lambda: k.call(Func.Inverse, [m], k.matrix(3)), "MatMul": lambda: m * mt,
             "MatCompMul": lambda: k.call(Func.MatCompMul, [m, mt], k.matrix(3)), "OuterProduct": lambda: k.call(Func.OuterProduct, [c[0], c[1]], k.matrix(3))}[n]()
        cols = [r.extract(j) for j in range(3)]
        return [k.vec(k.f324, q.x, q.y, q.z, k.f(0.0)) for q in cols]
    got = _run(device, _build(cases.MAT3, 3, "f324", "f324", emit_m, rows_per_item=3), [d["ga"], d["gb"], d["gc"]], len(cases.MAT3), np.float32, rows_per_item=3)
    for i, n in enumerate(cases.MAT3):
        exact = n in cases.EXACT_F
        _check_float("Mat3." + n, got[i], g["m3_" + n], 0 if exact else 1e-5, 0 if exact else 1e-5, lanes=3)


@pytest.mark.gpu
def test_integer_builtins_are_bit_identical(device):
    from luisa_compute_rs_b200.ir import Func
    g, d = np.load(GOLDEN), cases.inputs()
    ops = {"Add": lambda a, b: a + b, "Sub": lambda a, b: a - b, "Mul": lambda a, b: a * b, "Div": lambda a, b: a / b, "Rem": lambda a, b: a % b,
           "BitAnd": lambda a, b: a & b, "BitOr": lambda a, b: a | b, "BitXor": lambda a, b: a ^ b, "BitNot": lambda a, b: ~a, "Neg": lambda a, b: -a,
           "Shl": lambda a, b: a << b, "Shr": lambda a, b: a >> b}
    unary = {"PopCount", "Clz", "Ctz", "Reverse", "Abs"}

    def emit(ty):
        def e(k, n, v):
            a, b, s = v
            if n in ("Shl", "Shr"):
                return ops[n](a, s)
            if n in ops:
                return ops[n](a, b)
            return k.call(_f(Func, n), [a] if n in unary else [a, b], getattr(k, ty))
        return e
    names = [n for n in cases.U4 if n != "Reverse"]
    got = _run(device, _build(names, 3, "u324", "u324", emit("u324")), [d["ua"], d["ub"], d["us"]], len(names), np.uint32)
    for i, n in enumerate(names):
        # clz(0) / ctz(0) are __builtin_clz(0) / __builtin_ctz(0) in the reference (cpu_prelude.h): undefined — g++ returned 31 / 0 here,
        # clang with lzcnt / tzcnt 32.  The device returns 32 (CUDA __clz / __ffs - 1 semantics); zeros are left out of the comparison.
        keep = d["ua"] != 0 if n in ("Clz", "Ctz") else np.ones_like(d["ua"], bool)
        assert np.array_equal(got[i][keep], g["u4_" + n][keep]), "u32 %s: %r vs %r" % (n, got[i][:2], g["u4_" + n][:2])
        if n in ("Clz", "Ctz"):
            assert (got[i][~keep] == 32).all()
    got = _run(device, _build(cases.I4, 3, "i324", "i324", emit("i324")), [d["ia"], d["ib"], d["is_"]], len(cases.I4), np.int32)
    for i, n in enumerate(cases.I4):
        assert np.array_equal(got[i], g["i4_" + n]), "i32 %s: %r vs %r" % (n, got[i][:2], g["i4_" + n][:2])


@pytest.mark.gpu
def test_casts_and_predicates_are_bit_identical(device):
    from luisa_compute_rs_b200.ir import Func
    g, d = np.load(GOLDEN), cases.inputs()

    def emit(k, n, v):
        a, c = v
        one, zero = k.vec(k.u324, k.u(1)), k.vec(k.u324, k.u(0))
        if n == "IsNan":
            return a.is_nan().select(one, zero)
        if n == "IsInf":
            return a.unary(Func.IsInf, k.bool4).select(one, zero)
        if n == "CastU32":
            return c.abs().cast(k.u324)
        if n == "CastI32":
            return c.cast(k.i324).bitcast(k.u324)
        return a.bitcast(k.u324)
    got = _run(device, _build(cases.F4_TO_U4, 2, "f324", "u324", emit), [d["fa"], d["fcast"]], len(cases.F4_TO_U4), np.uint32)
    for i, n in enumerate(cases.F4_TO_U4):
        want = g["fu_" + n]   # CastU32 is taken of |x| on both sides: negative float -> unsigned is undefined in C++
        assert np.array_equal(got[i], want), "%s: %r vs %r" % (n, got[i][:2], want[:2])


@pytest.mark.gpu
def test_reverse_is_bit_reversal(device):
    """Documented divergence.  The reference `cpu` device implements Func::Reverse as a BYTE swap (cpu_prelude.h: `__brev` =
    __builtin_bswap32 — the golden vector), its CUDA backend and the DSL's documentation as BIT reversal (CUDA's __brev).  This
    device reverses bits, i.e. agrees with the reference's GPU backend; DESIGN.md §4.4 records the difference."""
    from luisa_compute_rs_b200.ir import Func
    g, d = np.load(GOLDEN), cases.inputs()
    got = _run(device, _build(["Reverse"], 1, "u324", "u324", lambda k, n, v: k.call(Func.Reverse, [v[0]], k.u324)), [d["ua"]], 1, np.uint32)[0]
    bits = np.array([int(format(int(x), "032b")[::-1], 2) for x in d["ua"].reshape(-1)], np.uint32).reshape(-1, 4)
    assert np.array_equal(got, bits)
    assert np.array_equal(g["u4_Reverse"], d["ua"].byteswap())


# ---- texel conversions ------------------------------------------------------------------------------------------------------------------
def test_pixel_conversion_vectors_are_what_the_compiled_reference_returns():
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import make_device_math_golden as mk
    g = np.load(os.path.join(os.path.dirname(GOLDEN), "pixel_conversion_reference.npz"))
    assert sorted(g.files) == sorted(["px_%s_%s" % (n, k) for n in mk.PIXEL_CASES for k in ("raw", "back")])
    v = mk.pixel_values()
    # by hand: unorm8 is roundf(x * 255) clamped (cpu_texture.h:74-86): halves round away from zero, out-of-range values clamp
    raw = g["px_Rgba8Unorm_raw"].reshape(-1)
    flat = v.reshape(-1)
    for x, want in ((0.5 / 255, 1), (1.5 / 255, 2), (2.5 / 255, 3), (-0.5, 0), (1.5, 255), (1.0, 255)):
        i = int(np.argwhere(flat == np.float32(x))[0, 0])
        assert raw[i] == want, (x, raw[i])
    if os.path.exists(REF_LIB):
        fresh = mk.compute_pixels(REF_LIB)
        for k in g.files:
            assert np.array_equal(fresh[k].view(np.uint8), g[k].view(np.uint8)), k


@pytest.mark.gpu
def test_texel_write_and_read_conversions_are_bit_identical(device):
    """Texture2dWrite of a float4 then Texture2dRead, per storage: the stored texels (downloaded) and the values read back equal what the
    reference `cpu` device's cpu_texture.h stores and returns — unorm8 / unorm16 rounding and clamping, binary16 rounding and overflow."""
    from luisa_compute_rs_b200 import ir
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import make_device_math_golden as mk
    g = np.load(os.path.join(os.path.dirname(GOLDEN), "pixel_conversion_reference.npz"))
    v = mk.pixel_values()
    k = ir.KernelBuilder(block_size=(32, 1, 1))
    img, vals, out = k.arg_tex2d(k.f324), k.arg_buffer(k.f324), k.arg_buffer(k.f324)

    def body():
        i = k.dispatch_id().x
        c = k.vec(k.u322, i, k.u(0))
        img.tex_write(c, vals.read(i))
        out.write(i, img.tex_read(c))
    k.body(body)
    k.finish()
    sh = device.create_shader(C.addressof(k.km), keep=k)
    vb = device.create_buffer_from_array(v)
    for name in mk.PIXEL_CASES:
        tex = device.create_tex2d(name, mk.PIXEL_W, 1)
        ob = device.create_buffer(mk.PIXEL_W, 16, 16)
        sh.dispatch((mk.PIXEL_W,), tex, vb, ob)
        back = ob.view().to_numpy(np.float32).reshape(-1, 4)
        raw = tex.to_numpy().reshape(g["px_%s_raw" % name].shape)
        assert np.array_equal(raw.view(np.uint8), g["px_%s_raw" % name].view(np.uint8)), "%s: stored texels differ at %r" % (name, np.argwhere(raw != g["px_%s_raw" % name])[:4])
        assert np.array_equal(back.view(np.uint32), g["px_%s_back" % name].view(np.uint32)), "%s: values read back differ" % name
        tex.destroy(); ob.destroy()
    vb.destroy(); sh.destroy()


@pytest.mark.gpu
def test_bindless_texture3d_sampling_is_bit_identical_to_the_reference_header(device):
    """BindlessTexture3dSample over 8 slots = point / trilinear x edge / repeat / mirror / zero of one Float4 volume against
    lc_texture_3d_sample of the compiled cpu_texture.h (tests/golden/texture_sample_reference.npz, keys tex3d_*)."""
    from luisa_compute_rs_b200 import ir
    from luisa_compute_rs_b200.ir import Func
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import make_device_math_golden as mk
    g = np.load(os.path.join(os.path.dirname(GOLDEN), "texture_sample_reference.npz"))
    vol, uvw = mk.texture3d_inputs()
    n = uvw.shape[0]
    k = ir.KernelBuilder(block_size=(64, 1, 1))
    heap, coords, out = k.arg_bindless(), k.arg_buffer(k.f324), k.arg_buffer(k.f324)

    def body():
        i = k.dispatch_id().x
        c = coords.read(i / k.u(8))
        out.write(i, k.call(Func.BindlessTexture3dSample, [heap, i % k.u(8), k.vec(k.f323, c.x, c.y, c.z)], k.f324))
    k.body(body)
    k.finish()
    tex = device.create_tex3d("Rgba32f", vol.shape[2], vol.shape[1], vol.shape[0]); tex.copy_from(vol)
    heap_arr = device.create_bindless_array(8)
    for filt in (0, 1):
        for address in range(4):
            heap_arr.emplace_tex3d_async(filt * 4 + address, tex, filt, address)
    heap_arr.update()
    cb = device.create_buffer_from_array(uvw); ob = device.create_buffer(n * 8, 16, 16)
    sh = device.create_shader(C.addressof(k.km), keep=k)
    sh.dispatch((n * 8,), heap_arr, cb, ob)
    got = ob.view().to_numpy(np.float32).reshape(n, 8, 4)
    for filt in (0, 1):
        for address in range(4):
            want = g["tex3d_%d_%d" % (filt, address)]
            x = got[:, filt * 4 + address]
            assert np.array_equal(x.view(np.uint32), want.view(np.uint32)), "filter %d address %d: %d of %d differ, max %g" % (filt, address, (x != want).any(1).sum(), n, np.abs(x - want).max())
    for r in (sh, cb, ob, heap_arr, tex):
        r.destroy()
