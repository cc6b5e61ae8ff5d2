// Test infrastructure (development container only): the reference CPU backend's kernel-side library, compiled.
//
// A kernel of the reference `cpu` device is C++ text: a few `using` lines + cpu_libm_def.h + cpu_kernel_defs.h + cpu_prelude.h +
// device_math.h + cpu_resource.h + cpu_texture.h + the generated body (cpu/codegen/cpp.rs:2041-2095), handed to clang++ -O3
// (cpu/shader.rs:44-67).  oracle/Makefile `ref_device_math` compiles exactly that preamble — the headers where they lie, via
// -include, in that order — followed by this file, with g++ into oracle/_ref/libref_device_math.so.  The entry points below do
// nothing but call the `lc_*` function the code generator emits for an IR `Func` (cpp.rs:520-640) on arrays, so that
// tests/golden/device_math_reference.npz (tests/golden/make_device_math_golden.py) holds what the reference computes for every
// builtin the IR -> CUDA lowering implements.  Only __fp16 is mapped (-D__fp16=_Float16: g++ on x86 has no __fp16);
// -ffp-contract=off so that the vectors do not depend on the host's FMA units (the GPU side is compiled -fmad=false).
#define STR_EQ(a, b) (__builtin_strcmp((a), (b)) == 0)

namespace {
template<class F> int map1(const float *a, float *o, size_t n, F f) {
    for (size_t i = 0; i < n; i++) { lc_float4 r = f(lc_make_float4(a[4 * i], a[4 * i + 1], a[4 * i + 2], a[4 * i + 3])); o[4 * i] = r.x; o[4 * i + 1] = r.y; o[4 * i + 2] = r.z; o[4 * i + 3] = r.w; }
    return 0;
}
inline lc_float4 ld4(const float *p, size_t i) { return lc_make_float4(p[4 * i], p[4 * i + 1], p[4 * i + 2], p[4 * i + 3]); }
inline lc_float3 ld3(const float *p, size_t i) { return lc_make_float3(p[4 * i], p[4 * i + 1], p[4 * i + 2]); }
inline void st4(float *p, size_t i, lc_float4 r) { p[4 * i] = r.x; p[4 * i + 1] = r.y; p[4 * i + 2] = r.z; p[4 * i + 3] = r.w; }
inline void st3(float *p, size_t i, lc_float3 r, float w) { p[4 * i] = r.x; p[4 * i + 1] = r.y; p[4 * i + 2] = r.z; p[4 * i + 3] = w; }
}// namespace

#define U1(NAME, FN) if (STR_EQ(name, NAME)) return map1(a, out, n, [](lc_float4 v) { return FN(v); });

// out[i] = F(a[i]) on float4
extern "C" int ref_f4_unary(const char *name, const float *a, float *out, size_t n) {
    U1("Abs", lc_abs) U1("Acos", lc_acos) U1("Acosh", lc_acosh) U1("Asin", lc_asin) U1("Asinh", lc_asinh) U1("Atan", lc_atan) U1("Atanh", lc_atanh)
    U1("Cos", lc_cos) U1("Cosh", lc_cosh) U1("Sin", lc_sin) U1("Sinh", lc_sinh) U1("Tan", lc_tan) U1("Tanh", lc_tanh)
    U1("Exp", lc_exp) U1("Exp2", lc_exp2) U1("Exp10", lc_exp10) U1("Log", lc_log) U1("Log2", lc_log2) U1("Log10", lc_log10)
    U1("Sqrt", lc_sqrt) U1("Rsqrt", lc_rsqrt) U1("Ceil", lc_ceil) U1("Floor", lc_floor) U1("Fract", lc_fract) U1("Trunc", lc_trunc) U1("Round", lc_round)
    U1("Saturate", lc_saturate) U1("Normalize", lc_normalize)
    if (STR_EQ(name, "Neg")) return map1(a, out, n, [](lc_float4 v) { return -v; });
    return -1;
}

// out[i] = F(a[i], b[i]) on float4
extern "C" int ref_f4_binary(const char *name, const float *a, const float *b, float *out, size_t n) {
    for (size_t i = 0; i < n; i++) {
        lc_float4 x = ld4(a, i), y = ld4(b, i), r;
        if (STR_EQ(name, "Atan2")) r = lc_atan2(x, y);
        else if (STR_EQ(name, "Powf")) r = lc_pow(x, y);
        else if (STR_EQ(name, "Copysign")) r = lc_copysign(x, y);
        else if (STR_EQ(name, "Min")) r = lc_min(x, y);
        else if (STR_EQ(name, "Max")) r = lc_max(x, y);
        else if (STR_EQ(name, "Step")) r = lc_step(x, y);
        else if (STR_EQ(name, "Add")) r = x + y;
        else if (STR_EQ(name, "Sub")) r = x - y;
        else if (STR_EQ(name, "Mul")) r = x * y;
        else if (STR_EQ(name, "Div")) r = x / y;
        else if (STR_EQ(name, "Rem")) r = lc_fmod(x, y);    // Func::Rem on floats: cpp.rs emits lc_fmod
        else return -1;
        st4(out, i, r);
    }
    return 0;
}

// out[i] = F(a[i], b[i], c[i]) on float4
extern "C" int ref_f4_ternary(const char *name, const float *a, const float *b, const float *c, float *out, size_t n) {
    for (size_t i = 0; i < n; i++) {
        lc_float4 x = ld4(a, i), y = ld4(b, i), z = ld4(c, i), r;
        if (STR_EQ(name, "Fma")) r = lc_fma(x, y, z);
        else if (STR_EQ(name, "Clamp")) r = lc_clamp(x, y, z);
        else if (STR_EQ(name, "Lerp")) r = lc_lerp(x, y, z);
        else if (STR_EQ(name, "SmoothStep")) r = lc_smoothstep(x, y, z);
        else return -1;
        st4(out, i, r);
    }
    return 0;
}

// float3 geometry on the xyz of float4 records; scalar results are broadcast to xyz, w = 0
extern "C" int ref_f3_geometry(const char *name, const float *a, const float *b, const float *c, float *out, size_t n) {
    for (size_t i = 0; i < n; i++) {
        lc_float3 x = ld3(a, i), y = ld3(b, i), z = ld3(c, i);
        if (STR_EQ(name, "Cross")) st3(out, i, lc_cross(x, y), 0.f);
        else if (STR_EQ(name, "Dot")) { float d = lc_dot(x, y); st3(out, i, lc_make_float3(d, d, d), 0.f); }
        else if (STR_EQ(name, "Length")) { float d = lc_length(x); st3(out, i, lc_make_float3(d, d, d), 0.f); }
        else if (STR_EQ(name, "LengthSquared")) { float d = lc_length_squared(x); st3(out, i, lc_make_float3(d, d, d), 0.f); }
        else if (STR_EQ(name, "Distance")) { float d = lc_distance(x, y); st3(out, i, lc_make_float3(d, d, d), 0.f); }
        else if (STR_EQ(name, "Normalize")) st3(out, i, lc_normalize(x), 0.f);
        else if (STR_EQ(name, "Faceforward")) st3(out, i, lc_faceforward(x, y, z), 0.f);
        else if (STR_EQ(name, "Reflect")) st3(out, i, lc_reflect(x, y), 0.f);
        else if (STR_EQ(name, "ReduceSum")) { float d = lc_reduce_sum(x); st3(out, i, lc_make_float3(d, d, d), 0.f); }
        else if (STR_EQ(name, "ReduceProd")) { float d = lc_reduce_prod(x); st3(out, i, lc_make_float3(d, d, d), 0.f); }
        else if (STR_EQ(name, "ReduceMin")) { float d = lc_reduce_min(x); st3(out, i, lc_make_float3(d, d, d), 0.f); }
        else if (STR_EQ(name, "ReduceMax")) { float d = lc_reduce_max(x); st3(out, i, lc_make_float3(d, d, d), 0.f); }
        else return -1;
    }
    return 0;
}

// 3x3 matrices: columns c0, c1, c2 are the xyz of a[i], b[i], c[i]; results are three float4 records per item (columns, w = 0)
extern "C" int ref_mat3(const char *name, const float *a, const float *b, const float *c, float *out, size_t n) {
    for (size_t i = 0; i < n; i++) {
        lc_float3x3 m = lc_make_float3x3(ld3(a, i), ld3(b, i), ld3(c, i)), r;
        if (STR_EQ(name, "Transpose")) r = lc_transpose(m);
        else if (STR_EQ(name, "Inverse")) r = lc_inverse(m);
        else if (STR_EQ(name, "MatMul")) r = m * lc_transpose(m);
        else if (STR_EQ(name, "MatCompMul")) r = lc_mat_comp_mul(m, lc_transpose(m));
        else if (STR_EQ(name, "OuterProduct")) r = lc_outer_product(ld3(a, i), ld3(b, i));
        else if (STR_EQ(name, "Determinant")) { float d = lc_determinant(m); r = lc_make_float3x3(lc_make_float3(d, d, d), lc_make_float3(d, d, d), lc_make_float3(d, d, d)); }
        else if (STR_EQ(name, "MatVec")) { lc_float3 v = m * ld3(c, i); r = lc_make_float3x3(v, v, v); }
        else return -1;
        for (int k = 0; k < 3; k++) st3(out, 3 * i + k, r[k], 0.f);
    }
    return 0;
}

// out[i] = F(a[i], b[i]) on uint4 (unary functions ignore b)
extern "C" int ref_u4(const char *name, const uint32_t *a, const uint32_t *b, uint32_t *out, size_t n) {
    for (size_t i = 0; i < n; i++) {
        lc_uint4 x = lc_make_uint4(a[4 * i], a[4 * i + 1], a[4 * i + 2], a[4 * i + 3]), y = lc_make_uint4(b[4 * i], b[4 * i + 1], b[4 * i + 2], b[4 * i + 3]), r;
        if (STR_EQ(name, "PopCount")) r = lc_popcount(x);
        else if (STR_EQ(name, "Clz")) r = lc_clz(x);
        else if (STR_EQ(name, "Ctz")) r = lc_ctz(x);
        else if (STR_EQ(name, "Reverse")) r = lc_reverse(x);
        else if (STR_EQ(name, "Min")) r = lc_min(x, y);
        else if (STR_EQ(name, "Max")) r = lc_max(x, y);
        else if (STR_EQ(name, "Add")) r = x + y;
        else if (STR_EQ(name, "Sub")) r = x - y;
        else if (STR_EQ(name, "Mul")) r = x * y;
        else if (STR_EQ(name, "Div")) r = x / y;
        else if (STR_EQ(name, "Rem")) r = x % y;
        else if (STR_EQ(name, "BitAnd")) r = x & y;
        else if (STR_EQ(name, "BitOr")) r = x | y;
        else if (STR_EQ(name, "BitXor")) r = x ^ y;
        else if (STR_EQ(name, "BitNot")) r = ~x;
        else if (STR_EQ(name, "Shl")) r = x << y;
        else if (STR_EQ(name, "Shr")) r = x >> y;
        else return -1;
        out[4 * i] = r.x; out[4 * i + 1] = r.y; out[4 * i + 2] = r.z; out[4 * i + 3] = r.w;
    }
    return 0;
}

// the same on int4 (two's complement in / out)
extern "C" int ref_i4(const char *name, const int32_t *a, const int32_t *b, int32_t *out, size_t n) {
    for (size_t i = 0; i < n; i++) {
        lc_int4 x = lc_make_int4(a[4 * i], a[4 * i + 1], a[4 * i + 2], a[4 * i + 3]), y = lc_make_int4(b[4 * i], b[4 * i + 1], b[4 * i + 2], b[4 * i + 3]), r;
        if (STR_EQ(name, "Abs")) r = lc_abs(x);
        else if (STR_EQ(name, "Neg")) r = -x;
        else if (STR_EQ(name, "Min")) r = lc_min(x, y);
        else if (STR_EQ(name, "Max")) r = lc_max(x, y);
        else if (STR_EQ(name, "Div")) r = x / y;
        else if (STR_EQ(name, "Rem")) r = x % y;
        else if (STR_EQ(name, "Shr")) r = x >> y;
        else return -1;
        out[4 * i] = r.x; out[4 * i + 1] = r.y; out[4 * i + 2] = r.z; out[4 * i + 3] = r.w;
    }
    return 0;
}

// casts and predicates: float4 -> uint4
extern "C" int ref_f4_to_u4(const char *name, const float *a, uint32_t *out, size_t n) {
    for (size_t i = 0; i < n; i++) {
        lc_float4 x = ld4(a, i);
        lc_uint4 r;
        if (STR_EQ(name, "IsNan")) { lc_bool4 p = lc_isnan(x); r = lc_make_uint4(p.x, p.y, p.z, p.w); }
        else if (STR_EQ(name, "IsInf")) { lc_bool4 p = lc_isinf(x); r = lc_make_uint4(p.x, p.y, p.z, p.w); }
        else if (STR_EQ(name, "CastU32")) r = lc_make_uint4(x);
        else if (STR_EQ(name, "CastI32")) { lc_int4 q = lc_make_int4(x); r = lc_make_uint4((uint32_t)q.x, (uint32_t)q.y, (uint32_t)q.z, (uint32_t)q.w); }
        else if (STR_EQ(name, "Bitcast")) r = lc_make_uint4(lc_bit_cast<lc_uint>(x.x), lc_bit_cast<lc_uint>(x.y), lc_bit_cast<lc_uint>(x.z), lc_bit_cast<lc_uint>(x.w));
        else return -1;
        out[4 * i] = r.x; out[4 * i + 1] = r.y; out[4 * i + 2] = r.z; out[4 * i + 3] = r.w;
    }
    return 0;
}

// cpu_texture.h: lc_texture_2d_sample on a one-level FLOAT4 image (what BindlessTexture2dSample / Texture sampling reach on the `cpu`
// device, cpu_texture.h:616-625 + 417-500).  img = h x w x 4 floats, uv = n x 2, out = n x 4.
extern "C" int ref_texture2d_sample(const float *img, uint32_t w, uint32_t h, const float *uv, size_t n, uint32_t filter, uint32_t address, float *out) {
    // the `cpu` device stores images in 4 x 4 pixel blocks (TextureView::_pixel2d); the row-major input is laid out through the
    // reference's own write2d, as its texture upload does
    const size_t blocks = (size_t)((w + 3) / 4) * ((h + 3) / 4);
    lc_float4 *store = new lc_float4[blocks * 16]();
    Texture tex{};
    tex.data = reinterpret_cast<uint8_t *>(store);
    tex.width = w; tex.height = h; tex.depth = 1;
    tex.storage = LC_PIXEL_STORAGE_FLOAT4; tex.dimension = 2; tex.mip_levels = 1; tex.pixel_stride_shift = 4;
    tex.mip_offsets[0] = 0;
    TextureView view = lc_texture_view(&tex, 0u);
    for (uint32_t y = 0; y < h; y++)
        for (uint32_t x = 0; x < w; x++) view.write2d<lc_float4, float>(lc_make_uint2(x, y), ld4(img, (size_t)y * w + x));
    LCSampler s{static_cast<LCSamplerAddress>(address), static_cast<LCSamplerFilter>(filter)};
    for (size_t i = 0; i < n; i++) st4(out, i, lc_texture_2d_sample(nullptr, &tex, s, lc_make_float2(uv[2 * i], uv[2 * i + 1])));
    delete[] store;
    return 0;
}

// cpu_texture.h pixel conversions: lc_texture2d_write<float4> followed by lc_texture2d_read<float4> on a w x 1 image of the given
// LCPixelStorage.  raw = the stored pixel (pixel_bytes each, in x order, read at the reference's own _pixel2d address), back = what a
// float read returns.  Covers the float -> unorm8 / unorm16 / half rounding and clamping rules (cpu_texture.h:58-135).
extern "C" int ref_texture2d_write_read(uint32_t storage, uint32_t pixel_shift, const float *values, uint32_t w, uint8_t *raw, float *back) {
    const size_t blocks = (size_t)((w + 3) / 4);
    uint8_t *store = new uint8_t[(blocks * 16) << pixel_shift]();
    Texture tex{};
    tex.data = store; tex.width = w; tex.height = 1; tex.depth = 1;
    tex.storage = (uint8_t)storage; tex.dimension = 2; tex.mip_levels = 1; tex.pixel_stride_shift = (uint8_t)pixel_shift;
    tex.mip_offsets[0] = 0;
    Texture2D arg{tex, 0};
    KernelFnArgs *k_args = nullptr;
    for (uint32_t x = 0; x < w; x++) lc_texture2d_write<lc_float4>(k_args, arg, lc_make_uint2(x, 0u), ld4(values, x));
    TextureView view = lc_texture_view(&tex, 0u);
    for (uint32_t x = 0; x < w; x++) {
        const uint8_t *p = view._pixel2d(lc_make_uint2(x, 0u));
        for (uint32_t b = 0; b < (1u << pixel_shift); b++) raw[((size_t)x << pixel_shift) + b] = p[b];
        st4(back, x, lc_texture2d_read<lc_float4>(k_args, arg, lc_make_uint2(x, 0u)));
    }
    delete[] store;
    return 0;
}

// cpu_texture.h: lc_texture_3d_sample (point / trilinear x edge / repeat / mirror / zero) on a one-level FLOAT4 volume stored in the
// reference's 4 x 4 x 4 blocks.  vol = d x h x w x 4 floats, uvw = n x 4 (xyz used), out = n x 4.
extern "C" int ref_texture3d_sample(const float *vol, uint32_t w, uint32_t h, uint32_t d, const float *uvw, size_t n, uint32_t filter, uint32_t address, float *out) {
    const size_t blocks = (size_t)((w + 3) / 4) * ((h + 3) / 4) * ((d + 3) / 4);
    lc_float4 *store = new lc_float4[blocks * 64]();
    Texture tex{};
    tex.data = reinterpret_cast<uint8_t *>(store);
    tex.width = w; tex.height = h; tex.depth = d;
    tex.storage = LC_PIXEL_STORAGE_FLOAT4; tex.dimension = 3; tex.mip_levels = 1; tex.pixel_stride_shift = 4;
    tex.mip_offsets[0] = 0;
    TextureView view = lc_texture_view(&tex, 0u);
    for (uint32_t z = 0; z < d; z++)
        for (uint32_t y = 0; y < h; y++)
            for (uint32_t x = 0; x < w; x++) view.write3d<lc_float4, float>(lc_make_uint3(x, y, z), ld4(vol, ((size_t)z * h + y) * w + x));
    LCSampler s{static_cast<LCSamplerAddress>(address), static_cast<LCSamplerFilter>(filter)};
    for (size_t i = 0; i < n; i++) st4(out, i, lc_texture_3d_sample(nullptr, &tex, s, ld3(uvw, i)));
    delete[] store;
    return 0;
}
