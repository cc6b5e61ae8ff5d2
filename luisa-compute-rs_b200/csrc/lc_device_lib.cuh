// lc_device_lib.cuh — the device library that IR-lowered kernels are compiled against (NVRTC, sm_100a).
//
// It plays the role the CPU backend's device_math.h + cpu_resource.h + cpu_texture.h play for its generated C++
// (luisa_compute_backend_impl/src/cpu/codegen/): vector / matrix / array value types with the IR's size and alignment
// rules (ir.rs:234-293), the math builtins behind ir::Func, and the resource accessors (buffers, textures, bindless
// arrays, the acceleration structure).  Where the operation order of a builtin is observable in fp32 it follows the
// reference's definition: dot = x*x' + y*y' + z*z' left to right (device_math.h:3562), cross as device_math.h:3553,
// length = sqrt(dot) (:3570), normalize = v * rsqrt(dot(v, v)) (:3588) with rsqrt(x) = 1 / sqrt(x) (cpu_prelude.h:7),
// clamp = min(max(v, lo), hi) (:3372), lerp = t * (b - a) + a (:3408), select(f, t, p) = p ? t : f (:2842).
// Ray queries call the per-thread traversal of trace_device.cuh, exactly where the reference's generated code calls
// lc_trace_closest / lc_trace_any through the Accel vtable (cpu_resource.h:288-294).
#pragma once
#include "trace_device.cuh"

// ---- value types ------------------------------------------------------------------------------------------------------
template <class T, int N> struct lc_vec;
#define LC_VEC_ALIGN(T, N) (sizeof(T) * (N) < 16 ? sizeof(T) * (N) : 16)
template <class T> struct alignas(LC_VEC_ALIGN(T, 2)) lc_vec<T, 2> {
    T x, y;
    lc_vec() = default;  // trivial: vectors may live in __shared__ arrays; T() / T{} still zero-initialise
    __device__ explicit lc_vec(T s) : x(s), y(s) {}
    __device__ lc_vec(T a, T b) : x(a), y(b) {}
    __device__ T &operator[](unsigned i) { return (&x)[i]; }
    __device__ const T &operator[](unsigned i) const { return (&x)[i]; }
};
template <class T> struct alignas(LC_VEC_ALIGN(T, 4)) lc_vec<T, 3> {
    T x, y, z;
    lc_vec() = default;
    __device__ explicit lc_vec(T s) : x(s), y(s), z(s) {}
    __device__ lc_vec(T a, T b, T c) : x(a), y(b), z(c) {}
    __device__ lc_vec(lc_vec<T, 2> a, T c) : x(a.x), y(a.y), z(c) {}
    __device__ lc_vec(T a, lc_vec<T, 2> b) : x(a), y(b.x), z(b.y) {}
    __device__ T &operator[](unsigned i) { return (&x)[i]; }
    __device__ const T &operator[](unsigned i) const { return (&x)[i]; }
};
template <class T> struct alignas(LC_VEC_ALIGN(T, 4)) lc_vec<T, 4> {
    T x, y, z, w;
    lc_vec() = default;
    __device__ explicit lc_vec(T s) : x(s), y(s), z(s), w(s) {}
    __device__ lc_vec(T a, T b, T c, T d) : x(a), y(b), z(c), w(d) {}
    __device__ lc_vec(lc_vec<T, 3> a, T d) : x(a.x), y(a.y), z(a.z), w(d) {}
    __device__ lc_vec(lc_vec<T, 2> a, lc_vec<T, 2> b) : x(a.x), y(a.y), z(b.x), w(b.y) {}
    __device__ lc_vec(lc_vec<T, 2> a, T c, T d) : x(a.x), y(a.y), z(c), w(d) {}
    __device__ T &operator[](unsigned i) { return (&x)[i]; }
    __device__ const T &operator[](unsigned i) const { return (&x)[i]; }
};

#define LC_VEC_TYPEDEFS(T, name)     \
    typedef T lc_##name;             \
    typedef lc_vec<T, 2> lc_##name##2; \
    typedef lc_vec<T, 3> lc_##name##3; \
    typedef lc_vec<T, 4> lc_##name##4;
LC_VEC_TYPEDEFS(bool, bool)
LC_VEC_TYPEDEFS(int8_t, char)
LC_VEC_TYPEDEFS(uint8_t, uchar)
LC_VEC_TYPEDEFS(int16_t, short)
LC_VEC_TYPEDEFS(uint16_t, ushort)
LC_VEC_TYPEDEFS(int32_t, int)
LC_VEC_TYPEDEFS(uint32_t, uint)
LC_VEC_TYPEDEFS(int64_t, long)
LC_VEC_TYPEDEFS(uint64_t, ulong)
LC_VEC_TYPEDEFS(float, float)
LC_VEC_TYPEDEFS(double, double)
// Float16 (ir.rs Primitive::Float16; the CPU backend's `half`, device_math.h): storage is IEEE binary16, arithmetic happens in fp32
// and every SSA value of the type is rounded back (round-to-nearest-even) when it is formed — for + - * / and sqrt that is the
// correctly rounded binary16 result.  Only the two conversions are defined: operators and math builtins reach it through float.
struct alignas(2) lc_f16 {
    unsigned short bits;
    lc_f16() = default;
    __device__ lc_f16(float f) { asm("cvt.rn.f16.f32 %0, %1;" : "=h"(bits) : "f"(f)); }
    __device__ operator float() const { float f; asm("cvt.f32.f16 %0, %1;" : "=f"(f) : "h"(bits)); return f; }
    __device__ static lc_f16 from_bits(unsigned short b) { lc_f16 h; h.bits = b; return h; }
};
LC_VEC_TYPEDEFS(lc_f16, half)
static_assert(sizeof(lc_half) == 2 && sizeof(lc_half2) == 4 && sizeof(lc_half3) == 8 && sizeof(lc_half4) == 8 && alignof(lc_half4) == 8, "IR vector layout rules for f16");
static_assert(sizeof(lc_float3) == 16 && alignof(lc_float3) == 16 && sizeof(lc_float2) == 8 && sizeof(lc_bool3) == 4 && sizeof(lc_double3) == 32 &&
                  alignof(lc_double3) == 16 && sizeof(lc_uint4) == 16,
              "IR vector layout rules (ir.rs:234-263)");

template <class T, size_t N> struct lc_array {
    T a[N];
    __device__ T &operator[](size_t i) { return a[i]; }
    __device__ const T &operator[](size_t i) const { return a[i]; }
};

template <int N> struct lc_mat {
    lc_vec<float, N> cols[N];
    lc_mat() = default;
    __device__ lc_vec<float, N> &operator[](unsigned i) { return cols[i]; }
    __device__ const lc_vec<float, N> &operator[](unsigned i) const { return cols[i]; }
    __device__ static lc_mat full(float s) { lc_mat m; for (int i = 0; i < N; i++) m.cols[i] = lc_vec<float, N>(s); return m; }
    __device__ lc_mat comp_mul(const lc_mat &o) const { lc_mat m; for (int i = 0; i < N; i++) for (int j = 0; j < N; j++) m.cols[i][j] = cols[i][j] * o.cols[i][j]; return m; }
};
typedef lc_mat<2> lc_float2x2;
typedef lc_mat<3> lc_float3x3;
typedef lc_mat<4> lc_float4x4;
static_assert(sizeof(lc_float2x2) == 16 && sizeof(lc_float3x3) == 48 && sizeof(lc_float4x4) == 64, "IR matrix layout rules (ir.rs:275-293)");
__device__ inline lc_float2x2 lc_make_mat(lc_float2 a, lc_float2 b) { lc_float2x2 m; m.cols[0] = a; m.cols[1] = b; return m; }
__device__ inline lc_float3x3 lc_make_mat(lc_float3 a, lc_float3 b, lc_float3 c) { lc_float3x3 m; m.cols[0] = a; m.cols[1] = b; m.cols[2] = c; return m; }
__device__ inline lc_float4x4 lc_make_mat(lc_float4 a, lc_float4 b, lc_float4 c, lc_float4 d) { lc_float4x4 m; m.cols[0] = a; m.cols[1] = b; m.cols[2] = c; m.cols[3] = d; return m; }

template <class T> struct lc_elem { typedef T type; enum { N = 1 }; };
template <class T, int M> struct lc_elem<lc_vec<T, M>> { typedef T type; enum { N = M }; };

template <class T> __device__ inline T lc_zero() { return T(); }
template <class T> struct lc_one_impl { __device__ static T get() { return T(1); } };
template <class T, int N> struct lc_one_impl<lc_vec<T, N>> { __device__ static lc_vec<T, N> get() { return lc_vec<T, N>(T(1)); } };
template <int N> struct lc_one_impl<lc_mat<N>> { __device__ static lc_mat<N> get() { return lc_mat<N>::full(1.0f); } };
template <class T, size_t N> struct lc_one_impl<lc_array<T, N>> { __device__ static lc_array<T, N> get() { lc_array<T, N> r; for (size_t i = 0; i < N; i++) r.a[i] = lc_one_impl<T>::get(); return r; } };
template <class T> __device__ inline T lc_one() { return lc_one_impl<T>::get(); }

template <class D, class S> __device__ inline D lc_bit_cast(const S &s) {
    static_assert(sizeof(D) == sizeof(S), "bitcast between types of different size");
    D d; memcpy(&d, &s, sizeof(D)); return d;
}

// ---- operators --------------------------------------------------------------------------------------------------------
#define LC_LOOP _Pragma("unroll") for (int i = 0; i < N; i++)
#define LC_BINOP(op)                                                                                                                            \
    template <class T, int N> __device__ inline lc_vec<T, N> operator op(lc_vec<T, N> a, lc_vec<T, N> b) { lc_vec<T, N> r; LC_LOOP r[i] = a[i] op b[i]; return r; } \
    template <class T, int N> __device__ inline lc_vec<T, N> operator op(lc_vec<T, N> a, T b) { lc_vec<T, N> r; LC_LOOP r[i] = a[i] op b; return r; }             \
    template <class T, int N> __device__ inline lc_vec<T, N> operator op(T a, lc_vec<T, N> b) { lc_vec<T, N> r; LC_LOOP r[i] = a op b[i]; return r; }
LC_BINOP(+) LC_BINOP(-) LC_BINOP(*) LC_BINOP(/) LC_BINOP(%) LC_BINOP(&) LC_BINOP(|) LC_BINOP(^) LC_BINOP(<<) LC_BINOP(>>)
#define LC_CMPOP(op)                                                                                                                               \
    template <class T, int N> __device__ inline lc_vec<bool, N> operator op(lc_vec<T, N> a, lc_vec<T, N> b) { lc_vec<bool, N> r; LC_LOOP r[i] = a[i] op b[i]; return r; } \
    template <class T, int N> __device__ inline lc_vec<bool, N> operator op(lc_vec<T, N> a, T b) { lc_vec<bool, N> r; LC_LOOP r[i] = a[i] op b; return r; }             \
    template <class T, int N> __device__ inline lc_vec<bool, N> operator op(T a, lc_vec<T, N> b) { lc_vec<bool, N> r; LC_LOOP r[i] = a op b[i]; return r; }
LC_CMPOP(==) LC_CMPOP(!=) LC_CMPOP(<) LC_CMPOP(<=) LC_CMPOP(>) LC_CMPOP(>=)
template <class T, int N> __device__ inline lc_vec<T, N> operator-(lc_vec<T, N> a) { lc_vec<T, N> r; LC_LOOP r[i] = -a[i]; return r; }
template <class T, int N> __device__ inline lc_vec<T, N> operator~(lc_vec<T, N> a) { lc_vec<T, N> r; LC_LOOP r[i] = ~a[i]; return r; }
template <int N> __device__ inline lc_vec<bool, N> operator!(lc_vec<bool, N> a) { lc_vec<bool, N> r; LC_LOOP r[i] = !a[i]; return r; }
// fmodf-free remainder for floats is not an IR operation; Rem on floats maps to fmod like the C++ backends
__device__ inline float operator_rem(float a, float b) { return fmodf(a, b); }

// matrices (column-major, device_math.h mat section)
template <int N> __device__ inline lc_vec<float, N> operator*(const lc_mat<N> &m, lc_vec<float, N> v) {
    lc_vec<float, N> r = m.cols[0] * v[0];
    for (int i = 1; i < N; i++) r = r + m.cols[i] * v[i];
    return r;
}
template <int N> __device__ inline lc_mat<N> operator*(const lc_mat<N> &a, const lc_mat<N> &b) { lc_mat<N> r; for (int i = 0; i < N; i++) r.cols[i] = a * b.cols[i]; return r; }
template <int N> __device__ inline lc_mat<N> operator*(const lc_mat<N> &a, float s) { lc_mat<N> r; for (int i = 0; i < N; i++) r.cols[i] = a.cols[i] * s; return r; }
template <int N> __device__ inline lc_mat<N> operator*(float s, const lc_mat<N> &a) { return a * s; }
template <int N> __device__ inline lc_mat<N> operator/(const lc_mat<N> &a, float s) { lc_mat<N> r; for (int i = 0; i < N; i++) r.cols[i] = a.cols[i] / s; return r; }
template <int N> __device__ inline lc_mat<N> operator+(const lc_mat<N> &a, const lc_mat<N> &b) { lc_mat<N> r; for (int i = 0; i < N; i++) r.cols[i] = a.cols[i] + b.cols[i]; return r; }
template <int N> __device__ inline lc_mat<N> operator-(const lc_mat<N> &a, const lc_mat<N> &b) { lc_mat<N> r; for (int i = 0; i < N; i++) r.cols[i] = a.cols[i] - b.cols[i]; return r; }
template <int N> __device__ inline lc_mat<N> operator-(const lc_mat<N> &a) { lc_mat<N> r; for (int i = 0; i < N; i++) r.cols[i] = -a.cols[i]; return r; }
template <int N> __device__ inline lc_mat<N> lc_transpose(const lc_mat<N> &a) { lc_mat<N> r; for (int i = 0; i < N; i++) for (int j = 0; j < N; j++) r.cols[i][j] = a.cols[j][i]; return r; }
__device__ inline float lc_determinant(const lc_float2x2 &m) { return m[0][0] * m[1][1] - m[1][0] * m[0][1]; }
__device__ inline float lc_determinant(const lc_float3x3 &m) {
    return m[0].x * (m[1].y * m[2].z - m[2].y * m[1].z) - m[1].x * (m[0].y * m[2].z - m[2].y * m[0].z) + m[2].x * (m[0].y * m[1].z - m[1].y * m[0].z);
}
__device__ inline lc_float2x2 lc_inverse(const lc_float2x2 &m) {
    const float inv = 1.0f / lc_determinant(m);
    return lc_make_mat(lc_float2(m[1][1] * inv, -m[0][1] * inv), lc_float2(-m[1][0] * inv, m[0][0] * inv));
}
__device__ inline lc_float3x3 lc_inverse(const lc_float3x3 &m) {
    const float inv = 1.0f / lc_determinant(m);
    return lc_make_mat(lc_float3((m[1].y * m[2].z - m[2].y * m[1].z) * inv, (m[2].y * m[0].z - m[0].y * m[2].z) * inv, (m[0].y * m[1].z - m[1].y * m[0].z) * inv),
                       lc_float3((m[2].x * m[1].z - m[1].x * m[2].z) * inv, (m[0].x * m[2].z - m[2].x * m[0].z) * inv, (m[1].x * m[0].z - m[0].x * m[1].z) * inv),
                       lc_float3((m[1].x * m[2].y - m[2].x * m[1].y) * inv, (m[2].x * m[0].y - m[0].x * m[2].y) * inv, (m[0].x * m[1].y - m[1].x * m[0].y) * inv));
}
__device__ inline float lc_determinant(const lc_float4x4 &m) {
    const float c00 = m[2].z * m[3].w - m[3].z * m[2].w, c02 = m[1].z * m[3].w - m[3].z * m[1].w, c03 = m[1].z * m[2].w - m[2].z * m[1].w;
    const float c04 = m[2].y * m[3].w - m[3].y * m[2].w, c06 = m[1].y * m[3].w - m[3].y * m[1].w, c07 = m[1].y * m[2].w - m[2].y * m[1].w;
    const float c08 = m[2].y * m[3].z - m[3].y * m[2].z, c10 = m[1].y * m[3].z - m[3].y * m[1].z, c11 = m[1].y * m[2].z - m[2].y * m[1].z;
    const float f0 = m[1].x, f1 = m[2].x, f2 = m[3].x;
    (void)f0; (void)f1; (void)f2;
    const float a0 = +(m[1].y * c00 - m[2].y * c02 + m[3].y * c03), a1 = -(m[1].x * c00 - m[2].x * c02 + m[3].x * c03);
    const float a2 = +(m[1].x * c04 - m[2].x * c06 + m[3].x * c07), a3 = -(m[1].x * c08 - m[2].x * c10 + m[3].x * c11);
    return m[0].x * a0 + m[0].y * a1 + m[0].z * a2 + m[0].w * a3;
}
__device__ inline lc_float4x4 lc_inverse(const lc_float4x4 &m) {  // cofactor expansion
    lc_float4x4 r;
    float a[16], o[16];
    for (int c = 0; c < 4; c++) for (int k = 0; k < 4; k++) a[c * 4 + k] = m[c][k];
    o[0] = a[5] * a[10] * a[15] - a[5] * a[11] * a[14] - a[9] * a[6] * a[15] + a[9] * a[7] * a[14] + a[13] * a[6] * a[11] - a[13] * a[7] * a[10];
    o[4] = -a[4] * a[10] * a[15] + a[4] * a[11] * a[14] + a[8] * a[6] * a[15] - a[8] * a[7] * a[14] - a[12] * a[6] * a[11] + a[12] * a[7] * a[10];
    o[8] = a[4] * a[9] * a[15] - a[4] * a[11] * a[13] - a[8] * a[5] * a[15] + a[8] * a[7] * a[13] + a[12] * a[5] * a[11] - a[12] * a[7] * a[9];
    o[12] = -a[4] * a[9] * a[14] + a[4] * a[10] * a[13] + a[8] * a[5] * a[14] - a[8] * a[6] * a[13] - a[12] * a[5] * a[10] + a[12] * a[6] * a[9];
    o[1] = -a[1] * a[10] * a[15] + a[1] * a[11] * a[14] + a[9] * a[2] * a[15] - a[9] * a[3] * a[14] - a[13] * a[2] * a[11] + a[13] * a[3] * a[10];
    o[5] = a[0] * a[10] * a[15] - a[0] * a[11] * a[14] - a[8] * a[2] * a[15] + a[8] * a[3] * a[14] + a[12] * a[2] * a[11] - a[12] * a[3] * a[10];
    o[9] = -a[0] * a[9] * a[15] + a[0] * a[11] * a[13] + a[8] * a[1] * a[15] - a[8] * a[3] * a[13] - a[12] * a[1] * a[11] + a[12] * a[3] * a[9];
    o[13] = a[0] * a[9] * a[14] - a[0] * a[10] * a[13] - a[8] * a[1] * a[14] + a[8] * a[2] * a[13] + a[12] * a[1] * a[10] - a[12] * a[2] * a[9];
    o[2] = a[1] * a[6] * a[15] - a[1] * a[7] * a[14] - a[5] * a[2] * a[15] + a[5] * a[3] * a[14] + a[13] * a[2] * a[7] - a[13] * a[3] * a[6];
    o[6] = -a[0] * a[6] * a[15] + a[0] * a[7] * a[14] + a[4] * a[2] * a[15] - a[4] * a[3] * a[14] - a[12] * a[2] * a[7] + a[12] * a[3] * a[6];
    o[10] = a[0] * a[5] * a[15] - a[0] * a[7] * a[13] - a[4] * a[1] * a[15] + a[4] * a[3] * a[13] + a[12] * a[1] * a[7] - a[12] * a[3] * a[5];
    o[14] = -a[0] * a[5] * a[14] + a[0] * a[6] * a[13] + a[4] * a[1] * a[14] - a[4] * a[2] * a[13] - a[12] * a[1] * a[6] + a[12] * a[2] * a[5];
    o[3] = -a[1] * a[6] * a[11] + a[1] * a[7] * a[10] + a[5] * a[2] * a[11] - a[5] * a[3] * a[10] - a[9] * a[2] * a[7] + a[9] * a[3] * a[6];
    o[7] = a[0] * a[6] * a[11] - a[0] * a[7] * a[10] - a[4] * a[2] * a[11] + a[4] * a[3] * a[10] + a[8] * a[2] * a[7] - a[8] * a[3] * a[6];
    o[11] = -a[0] * a[5] * a[11] + a[0] * a[7] * a[9] + a[4] * a[1] * a[11] - a[4] * a[3] * a[9] - a[8] * a[1] * a[7] + a[8] * a[3] * a[5];
    o[15] = a[0] * a[5] * a[10] - a[0] * a[6] * a[9] - a[4] * a[1] * a[10] + a[4] * a[2] * a[9] + a[8] * a[1] * a[6] - a[8] * a[2] * a[5];
    const float inv = 1.0f / (a[0] * o[0] + a[1] * o[4] + a[2] * o[8] + a[3] * o[12]);
    for (int c = 0; c < 4; c++) for (int k = 0; k < 4; k++) r[c][k] = o[c * 4 + k] * inv;
    return r;
}

// ---- builtins ---------------------------------------------------------------------------------------------------------
template <class T> __device__ inline T lc_select(T f, T t, bool p) { return p ? t : f; }
template <class T, int N> __device__ inline lc_vec<T, N> lc_select(lc_vec<T, N> f, lc_vec<T, N> t, lc_vec<bool, N> p) { lc_vec<T, N> r; LC_LOOP r[i] = p[i] ? t[i] : f[i]; return r; }
template <int N> __device__ inline bool lc_any(lc_vec<bool, N> v) { bool r = false; LC_LOOP r = r || v[i]; return r; }
template <int N> __device__ inline bool lc_all(lc_vec<bool, N> v) { bool r = true; LC_LOOP r = r && v[i]; return r; }
__device__ inline bool lc_any(bool v) { return v; }
__device__ inline bool lc_all(bool v) { return v; }

#define LC_UNARY_VEC(name) template <class T, int N> __device__ inline lc_vec<T, N> name(lc_vec<T, N> v) { lc_vec<T, N> r; LC_LOOP r[i] = name(v[i]); return r; }
#define LC_BINARY_VEC(name)                                                                                                                   \
    template <class T, int N> __device__ inline lc_vec<T, N> name(lc_vec<T, N> a, lc_vec<T, N> b) { lc_vec<T, N> r; LC_LOOP r[i] = name(a[i], b[i]); return r; } \
    template <class T, int N> __device__ inline lc_vec<T, N> name(lc_vec<T, N> a, T b) { lc_vec<T, N> r; LC_LOOP r[i] = name(a[i], b); return r; }             \
    template <class T, int N> __device__ inline lc_vec<T, N> name(T a, lc_vec<T, N> b) { lc_vec<T, N> r; LC_LOOP r[i] = name(a, b[i]); return r; }
#define LC_TERNARY_VEC(name) \
    template <class T, int N> __device__ inline lc_vec<T, N> name(lc_vec<T, N> a, lc_vec<T, N> b, lc_vec<T, N> c) { lc_vec<T, N> r; LC_LOOP r[i] = name(a[i], b[i], c[i]); return r; }
#define LC_FLOAT_UNARY(name, ff, fd)                  \
    __device__ inline float name(float x) { return ff(x); }  \
    __device__ inline double name(double x) { return fd(x); } \
    LC_UNARY_VEC(name)
LC_FLOAT_UNARY(lc_acos, acosf, acos) LC_FLOAT_UNARY(lc_acosh, acoshf, acosh) LC_FLOAT_UNARY(lc_asin, asinf, asin) LC_FLOAT_UNARY(lc_asinh, asinhf, asinh)
LC_FLOAT_UNARY(lc_atan, atanf, atan) LC_FLOAT_UNARY(lc_atanh, atanhf, atanh) LC_FLOAT_UNARY(lc_cos, cosf, cos) LC_FLOAT_UNARY(lc_cosh, coshf, cosh)
LC_FLOAT_UNARY(lc_sin, sinf, sin) LC_FLOAT_UNARY(lc_sinh, sinhf, sinh) LC_FLOAT_UNARY(lc_tan, tanf, tan) LC_FLOAT_UNARY(lc_tanh, tanhf, tanh)
LC_FLOAT_UNARY(lc_exp, expf, exp) LC_FLOAT_UNARY(lc_exp2, exp2f, exp2) LC_FLOAT_UNARY(lc_exp10, exp10f, exp10) LC_FLOAT_UNARY(lc_log, logf, log)
LC_FLOAT_UNARY(lc_log2, log2f, log2) LC_FLOAT_UNARY(lc_log10, log10f, log10) LC_FLOAT_UNARY(lc_sqrt, sqrtf, sqrt) LC_FLOAT_UNARY(lc_ceil, ceilf, ceil)
LC_FLOAT_UNARY(lc_floor, floorf, floor) LC_FLOAT_UNARY(lc_trunc, truncf, trunc) LC_FLOAT_UNARY(lc_round, roundf, round)
__device__ inline float lc_rsqrt(float x) { return 1.0f / sqrtf(x); }   // cpu_prelude.h:7 — not the approximate MUFU.RSQ
__device__ inline double lc_rsqrt(double x) { return 1.0 / sqrt(x); }
LC_UNARY_VEC(lc_rsqrt)
__device__ inline float lc_fract(float x) { return x - floorf(x); }
__device__ inline double lc_fract(double x) { return x - floor(x); }
LC_UNARY_VEC(lc_fract)
__device__ inline float lc_saturate(float x) { return fminf(fmaxf(x, 0.0f), 1.0f); }
LC_UNARY_VEC(lc_saturate)
__device__ inline float lc_abs(float x) { return fabsf(x); }
__device__ inline double lc_abs(double x) { return fabs(x); }
__device__ inline int8_t lc_abs(int8_t x) { return x < 0 ? -x : x; }
__device__ inline int16_t lc_abs(int16_t x) { return x < 0 ? -x : x; }
__device__ inline int32_t lc_abs(int32_t x) { return x < 0 ? -x : x; }
__device__ inline int64_t lc_abs(int64_t x) { return x < 0 ? -x : x; }
LC_UNARY_VEC(lc_abs)
#define LC_MINMAX_INT(T) __device__ inline T lc_min(T a, T b) { return a < b ? a : b; } __device__ inline T lc_max(T a, T b) { return a > b ? a : b; }
LC_MINMAX_INT(int8_t) LC_MINMAX_INT(uint8_t) LC_MINMAX_INT(int16_t) LC_MINMAX_INT(uint16_t) LC_MINMAX_INT(int32_t) LC_MINMAX_INT(uint32_t) LC_MINMAX_INT(int64_t) LC_MINMAX_INT(uint64_t)
__device__ inline float lc_min(float a, float b) { return fminf(a, b); }
__device__ inline float lc_max(float a, float b) { return fmaxf(a, b); }
__device__ inline double lc_min(double a, double b) { return fmin(a, b); }
__device__ inline double lc_max(double a, double b) { return fmax(a, b); }
LC_BINARY_VEC(lc_min) LC_BINARY_VEC(lc_max)
__device__ inline float lc_atan2(float a, float b) { return atan2f(a, b); }
__device__ inline float lc_pow(float a, float b) { return powf(a, b); }
__device__ inline double lc_pow(double a, double b) { return pow(a, b); }
__device__ inline float lc_copysign(float a, float b) { return copysignf(a, b); }
__device__ inline float lc_step(float edge, float x) { return x < edge ? 0.0f : 1.0f; }
LC_BINARY_VEC(lc_atan2) LC_BINARY_VEC(lc_pow) LC_BINARY_VEC(lc_copysign) LC_BINARY_VEC(lc_step)
template <class T> __device__ inline T lc_powi_scalar(T x, int32_t n) { T r = T(1); bool neg = n < 0; uint32_t k = neg ? 0u - (uint32_t)n : (uint32_t)n; while (k) { if (k & 1u) r = r * x; x = x * x; k >>= 1; } return neg ? T(1) / r : r; }
__device__ inline float lc_powi(float x, int32_t n) { return lc_powi_scalar(x, n); }
template <int N> __device__ inline lc_vec<float, N> lc_powi(lc_vec<float, N> x, int32_t n) { lc_vec<float, N> r; LC_LOOP r[i] = lc_powi_scalar(x[i], n); return r; }
template <int N> __device__ inline lc_vec<float, N> lc_powi(lc_vec<float, N> x, lc_vec<int32_t, N> n) { lc_vec<float, N> r; LC_LOOP r[i] = lc_powi_scalar(x[i], n[i]); return r; }
__device__ inline float lc_fma(float a, float b, float c) { return fmaf(a, b, c); }
__device__ inline double lc_fma(double a, double b, double c) { return fma(a, b, c); }
template <class T> __device__ inline T lc_clamp_scalar(T v, T lo, T hi) { return lc_min(lc_max(v, lo), hi); }
#define LC_CLAMP(T) __device__ inline T lc_clamp(T v, T lo, T hi) { return lc_clamp_scalar(v, lo, hi); }
LC_CLAMP(int8_t) LC_CLAMP(uint8_t) LC_CLAMP(int16_t) LC_CLAMP(uint16_t) LC_CLAMP(int32_t) LC_CLAMP(uint32_t) LC_CLAMP(int64_t) LC_CLAMP(uint64_t) LC_CLAMP(float) LC_CLAMP(double)
__device__ inline float lc_lerp(float a, float b, float t) { return t * (b - a) + a; }
__device__ inline float lc_smoothstep(float e0, float e1, float x) { const float t = lc_clamp((x - e0) / (e1 - e0), 0.0f, 1.0f); return t * t * (3.0f - 2.0f * t); }
LC_TERNARY_VEC(lc_fma) LC_TERNARY_VEC(lc_clamp) LC_TERNARY_VEC(lc_lerp) LC_TERNARY_VEC(lc_smoothstep)
template <class T, int N> __device__ inline lc_vec<T, N> lc_clamp(lc_vec<T, N> v, T lo, T hi) { lc_vec<T, N> r; LC_LOOP r[i] = lc_clamp(v[i], lo, hi); return r; }
template <class T, int N> __device__ inline lc_vec<T, N> lc_lerp(lc_vec<T, N> a, lc_vec<T, N> b, T t) { lc_vec<T, N> r; LC_LOOP r[i] = lc_lerp(a[i], b[i], t); return r; }

__device__ inline bool lc_isnan(float x) { return (__float_as_uint(x) & 0x7fffffffu) > 0x7f800000u; }
__device__ inline bool lc_isinf(float x) { return (__float_as_uint(x) & 0x7fffffffu) == 0x7f800000u; }
template <int N> __device__ inline lc_vec<bool, N> lc_isnan(lc_vec<float, N> v) { lc_vec<bool, N> r; LC_LOOP r[i] = lc_isnan(v[i]); return r; }
template <int N> __device__ inline lc_vec<bool, N> lc_isinf(lc_vec<float, N> v) { lc_vec<bool, N> r; LC_LOOP r[i] = lc_isinf(v[i]); return r; }

__device__ inline uint32_t lc_popcount(uint32_t x) { return __popc(x); }
__device__ inline uint32_t lc_clz(uint32_t x) { return __clz(x); }
__device__ inline uint32_t lc_ctz(uint32_t x) { return x ? __ffs(x) - 1 : 32u; }
__device__ inline uint32_t lc_reverse(uint32_t x) { return __brev(x); }
__device__ inline uint64_t lc_popcount(uint64_t x) { return __popcll(x); }
__device__ inline uint64_t lc_clz(uint64_t x) { return __clzll(x); }
__device__ inline uint64_t lc_ctz(uint64_t x) { return x ? __ffsll(x) - 1 : 64ull; }
__device__ inline uint64_t lc_reverse(uint64_t x) { return __brevll(x); }
LC_UNARY_VEC(lc_popcount) LC_UNARY_VEC(lc_clz) LC_UNARY_VEC(lc_ctz) LC_UNARY_VEC(lc_reverse)
__device__ inline uint32_t lc_rotl(uint32_t x, uint32_t s) { s &= 31u; return (x << s) | (x >> ((32u - s) & 31u)); }
__device__ inline uint32_t lc_rotr(uint32_t x, uint32_t s) { s &= 31u; return (x >> s) | (x << ((32u - s) & 31u)); }
__device__ inline uint64_t lc_rotl(uint64_t x, uint64_t s) { s &= 63u; return (x << s) | (x >> ((64u - s) & 63u)); }
__device__ inline uint64_t lc_rotr(uint64_t x, uint64_t s) { s &= 63u; return (x >> s) | (x << ((64u - s) & 63u)); }
LC_BINARY_VEC(lc_rotl) LC_BINARY_VEC(lc_rotr)

template <class T, int N> __device__ inline T lc_reduce_sum(lc_vec<T, N> v) { T r = v[0]; for (int i = 1; i < N; i++) r = r + v[i]; return r; }
template <class T, int N> __device__ inline T lc_reduce_prod(lc_vec<T, N> v) { T r = v[0]; for (int i = 1; i < N; i++) r = r * v[i]; return r; }
template <class T, int N> __device__ inline T lc_reduce_min(lc_vec<T, N> v) { T r = v[0]; for (int i = 1; i < N; i++) r = lc_min(r, v[i]); return r; }
template <class T, int N> __device__ inline T lc_reduce_max(lc_vec<T, N> v) { T r = v[0]; for (int i = 1; i < N; i++) r = lc_max(r, v[i]); return r; }
template <class T, int N> __device__ inline T lc_dot(lc_vec<T, N> a, lc_vec<T, N> b) { T r = a[0] * b[0]; for (int i = 1; i < N; i++) r = r + a[i] * b[i]; return r; }
template <class T> __device__ inline lc_vec<T, 3> lc_cross(lc_vec<T, 3> u, lc_vec<T, 3> v) { return lc_vec<T, 3>(u.y * v.z - v.y * u.z, u.z * v.x - v.z * u.x, u.x * v.y - v.x * u.y); }
template <class T, int N> __device__ inline T lc_length_squared(lc_vec<T, N> v) { return lc_dot(v, v); }
template <class T, int N> __device__ inline T lc_length(lc_vec<T, N> v) { return lc_sqrt(lc_dot(v, v)); }
template <class T, int N> __device__ inline T lc_distance(lc_vec<T, N> a, lc_vec<T, N> b) { return lc_length(a - b); }
template <class T, int N> __device__ inline lc_vec<T, N> lc_normalize(lc_vec<T, N> v) { return v * lc_rsqrt(lc_dot(v, v)); }
__device__ inline lc_float3 lc_faceforward(lc_float3 n, lc_float3 i, lc_float3 n_ref) { return lc_dot(n_ref, i) < 0.0f ? n : -n; }
__device__ inline lc_float3 lc_reflect(lc_float3 v, lc_float3 n) { return v - 2.0f * lc_dot(v, n) * n; }
template <int N> __device__ inline lc_mat<N> lc_outer_product(lc_vec<float, N> a, lc_vec<float, N> b) { lc_mat<N> m; for (int i = 0; i < N; i++) m.cols[i] = a * b[i]; return m; }

__device__ inline float lc_fmod(float a, float b) { return fmodf(a, b); }
__device__ inline double lc_fmod(double a, double b) { return fmod(a, b); }
LC_BINARY_VEC(lc_fmod)

// ---- atomics (cpp.rs:1259-1310 lowers Func::Atomic* to lc_atomic_*; all return the old value) ------------------------------
template <class T> __device__ inline T lc_atomic_exchange(T *p, T v) { return atomicExch(p, v); }
template <class T> __device__ inline T lc_atomic_compare_exchange(T *p, T expected, T desired) { return atomicCAS(p, expected, desired); }
__device__ inline float lc_atomic_compare_exchange(float *p, float expected, float desired) {
    return __uint_as_float(atomicCAS(reinterpret_cast<unsigned int *>(p), __float_as_uint(expected), __float_as_uint(desired)));
}
__device__ inline int64_t lc_atomic_exchange(int64_t *p, int64_t v) { return (int64_t)atomicExch(reinterpret_cast<unsigned long long *>(p), (unsigned long long)v); }
__device__ inline int64_t lc_atomic_compare_exchange(int64_t *p, int64_t e, int64_t d) { return (int64_t)atomicCAS(reinterpret_cast<unsigned long long *>(p), (unsigned long long)e, (unsigned long long)d); }
template <class T> __device__ inline T lc_atomic_fetch_add(T *p, T v) { return atomicAdd(p, v); }
__device__ inline int64_t lc_atomic_fetch_add(int64_t *p, int64_t v) { return (int64_t)atomicAdd(reinterpret_cast<unsigned long long *>(p), (unsigned long long)v); }
template <class T> __device__ inline T lc_atomic_fetch_sub(T *p, T v) { return lc_atomic_fetch_add(p, (T)(T(0) - v)); }
__device__ inline float lc_atomic_fetch_sub(float *p, float v) { return atomicAdd(p, -v); }
template <class T> __device__ inline T lc_atomic_fetch_and(T *p, T v) { return atomicAnd(p, v); }
template <class T> __device__ inline T lc_atomic_fetch_or(T *p, T v) { return atomicOr(p, v); }
template <class T> __device__ inline T lc_atomic_fetch_xor(T *p, T v) { return atomicXor(p, v); }
template <class T> __device__ inline T lc_atomic_fetch_min(T *p, T v) { return atomicMin(p, v); }
template <class T> __device__ inline T lc_atomic_fetch_max(T *p, T v) { return atomicMax(p, v); }
__device__ inline float lc_atomic_fetch_min(float *p, float v) {
    float old = *p;
    for (;;) { const float seen = lc_atomic_compare_exchange(p, old, fminf(old, v)); if (__float_as_uint(seen) == __float_as_uint(old)) return old; old = seen; }
}
__device__ inline float lc_atomic_fetch_max(float *p, float v) {
    float old = *p;
    for (;;) { const float seen = lc_atomic_compare_exchange(p, old, fmaxf(old, v)); if (__float_as_uint(seen) == __float_as_uint(old)) return old; old = seen; }
}

// ---- warp intrinsics -----------------------------------------------------------------------------------------------------
// Every operation takes the mask `m` of the lanes taking part.  At the top level of a kernel body the lowering passes the mask of
// the warp's live lanes, computed once at kernel entry (lc_warp_mask): the *_sync primitives then wait for exactly those lanes, so
// the result does not depend on whether the hardware has reconverged after earlier divergent code (an If, the CAS loop of a float
// atomic).  Inside divergent constructs it passes __activemask(): the lanes that are there.
__device__ inline uint32_t lc_lane_id() { return (threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z)) & 31u; }
__device__ inline bool lc_warp_is_first_active_lane(uint32_t m) { return lc_lane_id() == (uint32_t)(__ffs(m) - 1); }
__device__ inline uint32_t lc_warp_first_active_lane(uint32_t m) { return (uint32_t)(__ffs(m) - 1); }
__device__ inline bool lc_warp_active_all(uint32_t m, bool v) { return __all_sync(m, v); }
__device__ inline bool lc_warp_active_any(uint32_t m, bool v) { return __any_sync(m, v); }
__device__ inline lc_uint4 lc_warp_active_bit_mask(uint32_t m, bool v) { return lc_uint4(__ballot_sync(m, v), 0u, 0u, 0u); }
__device__ inline uint32_t lc_warp_active_count_bits(uint32_t m, bool v) { return __popc(__ballot_sync(m, v)); }
__device__ inline uint32_t lc_warp_prefix_count_bits(uint32_t m, bool v) { return __popc(__ballot_sync(m, v) & ((1u << lc_lane_id()) - 1u)); }
template <class T> __device__ inline T lc_warp_read_lane_at(uint32_t m, T v, uint32_t lane) { return __shfl_sync(m, v, lane); }
template <class T> __device__ inline T lc_warp_read_first_lane(uint32_t m, T v) { return __shfl_sync(m, v, __ffs(m) - 1); }
template <class T, class Op> __device__ inline T lc_warp_reduce(uint32_t m, T v, Op op) {
    T r = v; bool have = false;
    for (uint32_t lanes = m; lanes; lanes &= lanes - 1) { const T o = __shfl_sync(m, v, __ffs(lanes) - 1); r = have ? op(r, o) : o; have = true; }
    return r;
}
template <class T, class Op> __device__ inline T lc_warp_prefix(uint32_t m, T v, T identity, Op op) {
    const uint32_t me = lc_lane_id();
    T r = identity;
    for (uint32_t lanes = m; lanes; lanes &= lanes - 1) { const uint32_t l = __ffs(lanes) - 1; const T o = __shfl_sync(m, v, l); if (l < me) r = op(r, o); }
    return r;
}
struct lc_op_add { template <class T> __device__ T operator()(T a, T b) const { return a + b; } };
struct lc_op_mul { template <class T> __device__ T operator()(T a, T b) const { return a * b; } };
struct lc_op_min { template <class T> __device__ T operator()(T a, T b) const { return lc_min(a, b); } };
struct lc_op_max { template <class T> __device__ T operator()(T a, T b) const { return lc_max(a, b); } };
struct lc_op_and { template <class T> __device__ T operator()(T a, T b) const { return a & b; } };
struct lc_op_or { template <class T> __device__ T operator()(T a, T b) const { return a | b; } };
struct lc_op_xor { template <class T> __device__ T operator()(T a, T b) const { return a ^ b; } };
template <class T> __device__ inline T lc_warp_active_sum(uint32_t m, T v) { return lc_warp_reduce(m, v, lc_op_add()); }
template <class T> __device__ inline T lc_warp_active_product(uint32_t m, T v) { return lc_warp_reduce(m, v, lc_op_mul()); }
template <class T> __device__ inline T lc_warp_active_min(uint32_t m, T v) { return lc_warp_reduce(m, v, lc_op_min()); }
template <class T> __device__ inline T lc_warp_active_max(uint32_t m, T v) { return lc_warp_reduce(m, v, lc_op_max()); }
template <class T> __device__ inline T lc_warp_active_bit_and(uint32_t m, T v) { return lc_warp_reduce(m, v, lc_op_and()); }
template <class T> __device__ inline T lc_warp_active_bit_or(uint32_t m, T v) { return lc_warp_reduce(m, v, lc_op_or()); }
template <class T> __device__ inline T lc_warp_active_bit_xor(uint32_t m, T v) { return lc_warp_reduce(m, v, lc_op_xor()); }
template <class T> __device__ inline bool lc_warp_active_all_equal(uint32_t m, T v) { return __all_sync(m, v == lc_warp_read_first_lane(m, v)); }
template <class T> __device__ inline T lc_warp_prefix_sum(uint32_t m, T v) { return lc_warp_prefix(m, v, T(0), lc_op_add()); }
template <class T> __device__ inline T lc_warp_prefix_product(uint32_t m, T v) { return lc_warp_prefix(m, v, T(1), lc_op_mul()); }

// casts between vectors (Func::Cast on vectors, cpp.rs:1218-1226)
template <class D, class S, int N> __device__ inline lc_vec<D, N> lc_vec_cast(lc_vec<S, N> v) { lc_vec<D, N> r; LC_LOOP r[i] = static_cast<D>(v[i]); return r; }

// ---- dispatch geometry ---------------------------------------------------------------------------------------------------
struct lc_launch { uint32_t dispatch_size[3]; uint32_t yield_min; unsigned long long *work_counter; unsigned long long work_items; };  // shader.h HostLaunch
static_assert(sizeof(lc_launch) == 32, "launch record is mirrored in shader.h");
__device__ inline lc_uint3 lc_thread_id() { return lc_uint3(threadIdx.x, threadIdx.y, threadIdx.z); }
__device__ inline lc_uint3 lc_block_id() { return lc_uint3(blockIdx.x, blockIdx.y, blockIdx.z); }
__device__ inline lc_uint3 lc_dispatch_id() { return lc_uint3(blockIdx.x * blockDim.x + threadIdx.x, blockIdx.y * blockDim.y + threadIdx.y, blockIdx.z * blockDim.z + threadIdx.z); }
// thread_id / block_id / dispatch_id of the DSL as the kernel body and its callables see them.  Direct lowering: the CUDA thread's own;
// wavefront lowering: those of the work item the lane is running (lc_wave_ids).
struct lc_ids_t { lc_uint3 thread, block, dispatch; };
// work item -> ids: items enumerate the grid the direct lowering would launch, block-major and x-fastest inside a block (the order of
// a CUDA block's threads), so consecutive items are neighbours in the dispatch.  false: the position lies outside dispatch_size.
template <uint32_t BX, uint32_t BY, uint32_t BZ> __device__ inline bool lc_wave_ids(const lc_launch &l, unsigned long long item, lc_ids_t &ids) {
    constexpr uint32_t kBlock = BX * BY * BZ;
    if (BY == 1 && BZ == 1 && l.dispatch_size[1] == 1 && l.dispatch_size[2] == 1 && l.work_items <= 0xffffffffull) {
        // one-dimensional dispatch (buffer-to-buffer kernels): the work item IS the dispatch id — no divisions on the per-ray path
        const uint32_t x = (uint32_t)item;
        ids.block = lc_uint3(x / BX, 0u, 0u);
        ids.thread = lc_uint3(x % BX, 0u, 0u);
        ids.dispatch = lc_uint3(x, 0u, 0u);
        return x < l.dispatch_size[0];
    }
    const uint32_t gx = (l.dispatch_size[0] + BX - 1) / BX, gy = (l.dispatch_size[1] + BY - 1) / BY;
    uint32_t t, bxy, bz;
    if (l.work_items <= 0xffffffffull) {  // 32-bit arithmetic whenever the padded grid fits (a 64-bit division costs ~100 instructions)
        const uint32_t b = (uint32_t)item / kBlock;
        t = (uint32_t)item - b * kBlock;
        bz = b / (gx * gy);
        bxy = b - bz * (gx * gy);
    } else {
        const unsigned long long b = item / kBlock;
        t = (uint32_t)(item - b * kBlock);
        const unsigned long long z = b / ((unsigned long long)gx * gy);
        bz = (uint32_t)z;
        bxy = (uint32_t)(b - z * ((unsigned long long)gx * gy));
    }
    ids.block = lc_uint3(bxy % gx, bxy / gx, bz);
    ids.thread = lc_uint3(t % BX, (t / BX) % BY, t / (BX * BY));
    ids.dispatch = lc_uint3(ids.block.x * BX + ids.thread.x, ids.block.y * BY + ids.thread.y, ids.block.z * BZ + ids.thread.z);
    return ids.dispatch.x < l.dispatch_size[0] && ids.dispatch.y < l.dispatch_size[1] && ids.dispatch.z < l.dispatch_size[2];
}
__device__ inline void lc_assume(bool) {}
__device__ inline void lc_trap(const char *what, int id) { printf("[lc_b200 kernel] %s (message %d) at block (%u,%u,%u) thread (%u,%u,%u)\n", what, id, blockIdx.x, blockIdx.y, blockIdx.z, threadIdx.x, threadIdx.y, threadIdx.z); __trap(); }
__device__ inline void lc_assert(bool c, int id) { if (!c) lc_trap("assertion failed", id); }

// ---- resources -------------------------------------------------------------------------------------------------------------
// Kernel parameter records written by the host at every ShaderDispatch (shader.cu: pack_arguments).
struct lc_buffer { uint8_t *ptr; uint64_t size; };                                     // BufferView of cpu_kernel_defs (data + byte size)
struct lc_texture { uint8_t *data; uint32_t width, height, depth; uint32_t storage; uint32_t sampler; uint32_t pad; }; // level-0 view, row-major texels; sampler = filter | address << 2
struct lc_bindless_slot { uint8_t *buffer; uint64_t buffer_size; lc_texture tex2d; lc_texture tex3d; };
struct lc_bindless { const lc_bindless_slot *slots; uint64_t count; };
struct lc_accel { lcb::AccelView view; lcb::InstanceRec *instances_rw; uint32_t *dirty; };

template <class T> __device__ inline T lc_buffer_read(const lc_buffer &b, uint64_t i) { return reinterpret_cast<const T *>(b.ptr)[i]; }
template <class T> __device__ inline void lc_buffer_write(const lc_buffer &b, uint64_t i, const T &v) { reinterpret_cast<T *>(b.ptr)[i] = v; }
template <class T> __device__ inline T &lc_buffer_ref(const lc_buffer &b, uint64_t i) { return reinterpret_cast<T *>(b.ptr)[i]; }
template <class T> __device__ inline uint64_t lc_buffer_size(const lc_buffer &b) { return b.size / sizeof(T); }
__device__ inline uint64_t lc_buffer_address(const lc_buffer &b) { return (uint64_t)b.ptr; }
template <class T> __device__ inline T lc_byte_buffer_read(const lc_buffer &b, uint64_t off) { T v; memcpy(&v, b.ptr + off, sizeof(T)); return v; }
template <class T> __device__ inline void lc_byte_buffer_write(const lc_buffer &b, uint64_t off, const T &v) { memcpy(b.ptr + off, &v, sizeof(T)); }
template <class T> __device__ inline T lc_bindless_buffer_read(const lc_bindless &a, uint32_t slot, uint64_t i) { return reinterpret_cast<const T *>(a.slots[slot].buffer)[i]; }
template <class T> __device__ inline void lc_bindless_buffer_write(const lc_bindless &a, uint32_t slot, uint64_t i, const T &v) { reinterpret_cast<T *>(a.slots[slot].buffer)[i] = v; }
template <class T> __device__ inline T lc_bindless_byte_buffer_read(const lc_bindless &a, uint32_t slot, uint64_t off) { T v; memcpy(&v, a.slots[slot].buffer + off, sizeof(T)); return v; }
__device__ inline uint64_t lc_bindless_buffer_size(const lc_bindless &a, uint32_t slot, uint64_t stride) { return a.slots[slot].buffer_size / stride; }
__device__ inline uint64_t lc_bindless_buffer_address(const lc_bindless &a, uint32_t slot) { return (uint64_t)a.slots[slot].buffer; }

// PixelStorage (api_types:366-383): BYTE1,2,4 SHORT1,2,4 INT1,2,4 HALF1,2,4 FLOAT1,2,4
__device__ inline uint32_t lc_storage_channels(uint32_t s) { const uint32_t k = s % 3u; return k == 0 ? 1u : (k == 1 ? 2u : 4u); }
__device__ inline uint32_t lc_storage_channel_bytes(uint32_t s) { const uint32_t g = s / 3u; return g == 0 ? 1u : (g == 1 || g == 3 ? 2u : 4u); }
__device__ inline float lc_half_bits_to_float(uint16_t h) { float f; asm("{ .reg .b16 t; mov.b16 t, %1; cvt.f32.f16 %0, t; }" : "=f"(f) : "h"(h)); return f; }
__device__ inline uint16_t lc_float_to_half_bits(float f) { uint16_t h; asm("{ .reg .b16 t; cvt.rn.f16.f32 t, %1; mov.b16 %0, t; }" : "=h"(h) : "f"(f)); return h; }
__device__ inline float lc_channel_to_float(const uint8_t *p, uint32_t group) {  // cpu_texture.h scalar_to_float
    switch (group) {
        case 0: return *p / 255.f;
        case 1: return *reinterpret_cast<const uint16_t *>(p) / 65535.f;
        case 3: return lc_half_bits_to_float(*reinterpret_cast<const uint16_t *>(p));
        case 4: return *reinterpret_cast<const float *>(p);
        default: return 0.f;  // INT storage read as float (cpu_texture.h:68-70)
    }
}
__device__ inline void lc_float_to_channel(uint8_t *p, uint32_t group, float x) {  // cpu_texture.h float_to_scalar
    switch (group) {
        case 0: *p = (uint8_t)lc_clamp(roundf(x * 255.f), 0.f, 255.f); break;
        case 1: *reinterpret_cast<uint16_t *>(p) = (uint16_t)lc_clamp(roundf(x * 65535.f), 0.f, 65535.f); break;
        case 3: *reinterpret_cast<uint16_t *>(p) = lc_float_to_half_bits(x); break;
        case 4: *reinterpret_cast<float *>(p) = x; break;
        default: *reinterpret_cast<uint32_t *>(p) = 0u; break;
    }
}
__device__ inline uint32_t lc_channel_to_uint(const uint8_t *p, uint32_t bytes) { return bytes == 1 ? *p : (bytes == 2 ? *reinterpret_cast<const uint16_t *>(p) : *reinterpret_cast<const uint32_t *>(p)); }
__device__ inline void lc_uint_to_channel(uint8_t *p, uint32_t bytes, uint32_t v) { if (bytes == 1) *p = (uint8_t)v; else if (bytes == 2) *reinterpret_cast<uint16_t *>(p) = (uint16_t)v; else *reinterpret_cast<uint32_t *>(p) = v; }
template <class E> struct lc_is_float { enum { value = 0 }; };
template <> struct lc_is_float<float> { enum { value = 1 }; };

template <class V> __device__ inline V lc_texel_read(const lc_texture &t, uint64_t texel) {
    typedef typename lc_elem<V>::type E;
    const uint32_t ch = lc_storage_channels(t.storage), cb = lc_storage_channel_bytes(t.storage);
    const uint8_t *p = t.data + texel * (uint64_t)(ch * cb);
    E c[4] = {E(0), E(0), E(0), E(0)};
    for (uint32_t k = 0; k < ch; k++) c[k] = lc_is_float<E>::value ? (E)lc_channel_to_float(p + k * cb, t.storage / 3u) : (E)lc_channel_to_uint(p + k * cb, cb);
    V v;
    memcpy(&v, c, sizeof(E) * lc_elem<V>::N);
    return v;
}
template <class V> __device__ inline void lc_texel_write(const lc_texture &t, uint64_t texel, const V &v) {
    typedef typename lc_elem<V>::type E;
    const uint32_t ch = lc_storage_channels(t.storage), cb = lc_storage_channel_bytes(t.storage);
    uint8_t *p = t.data + texel * (uint64_t)(ch * cb);
    E c[4] = {E(0), E(0), E(0), E(0)};
    memcpy(c, &v, sizeof(E) * lc_elem<V>::N);
    for (uint32_t k = 0; k < ch; k++) {
        if (lc_is_float<E>::value) lc_float_to_channel(p + k * cb, t.storage / 3u, (float)c[k]);
        else lc_uint_to_channel(p + k * cb, cb, (uint32_t)c[k]);
    }
}
// out-of-range coordinates read as zero and are not written (TextureView::read2d / write2d, cpu_texture.h:369-390)
template <class V> __device__ inline V lc_texture2d_read(const lc_texture &t, lc_uint2 uv) { return uv.x < t.width && uv.y < t.height ? lc_texel_read<V>(t, (uint64_t)uv.y * t.width + uv.x) : V(); }
template <class V> __device__ inline void lc_texture2d_write(const lc_texture &t, lc_uint2 uv, const V &v) { if (uv.x < t.width && uv.y < t.height) lc_texel_write<V>(t, (uint64_t)uv.y * t.width + uv.x, v); }
template <class V> __device__ inline V lc_texture3d_read(const lc_texture &t, lc_uint3 p) { return p.x < t.width && p.y < t.height && p.z < t.depth ? lc_texel_read<V>(t, ((uint64_t)p.z * t.height + p.y) * t.width + p.x) : V(); }
template <class V> __device__ inline void lc_texture3d_write(const lc_texture &t, lc_uint3 p, const V &v) { if (p.x < t.width && p.y < t.height && p.z < t.depth) lc_texel_write<V>(t, ((uint64_t)p.z * t.height + p.y) * t.width + p.x, v); }
__device__ inline lc_uint2 lc_texture2d_size(const lc_texture &t) { return lc_uint2(t.width, t.height); }
__device__ inline lc_uint3 lc_texture3d_size(const lc_texture &t) { return lc_uint3(t.width, t.height, t.depth); }
__device__ inline lc_float4 lc_bindless_texture2d_read(const lc_bindless &a, uint32_t slot, lc_uint2 uv) { return lc_texture2d_read<lc_float4>(a.slots[slot].tex2d, uv); }
__device__ inline lc_float4 lc_bindless_texture3d_read(const lc_bindless &a, uint32_t slot, lc_uint3 p) { return lc_texture3d_read<lc_float4>(a.slots[slot].tex3d, p); }
__device__ inline lc_uint2 lc_bindless_texture2d_size(const lc_bindless &a, uint32_t slot) { return lc_texture2d_size(a.slots[slot].tex2d); }
// Filtered sampling of bindless textures (texture_coord_point / texture_sample_point / texture_sample_linear, cpu_texture.h:417-500):
// address modes EDGE / REPEAT / MIRROR / ZERO, filters POINT and LINEAR_* (bilinear / trilinear in the texture's one level; the
// level / gradient arguments of the *Level / *Grad forms select nothing on a single-level texture).
__device__ inline float lc_sample_coord(uint32_t address, float uv, float s) {
    const float one_minus_epsilon = __uint_as_float(0x3f7fffffu);
    switch (address) {
        case 0: return lc_clamp(uv, 0.0f, one_minus_epsilon) * s;
        case 1: return lc_fract(uv) * s;
        case 2: { float m = fmodf(fabsf(uv), 2.0f); m = m < 1.0f ? m : 2.0f - m; return fminf(m, one_minus_epsilon) * s; }
        default: return (uv < 0.0f || uv >= 1.0f) ? 65536.0f : uv * s;
    }
}
__device__ inline lc_float4 lc_texture2d_sample(const lc_texture &t, lc_float2 uv) {
    const uint32_t filter = t.sampler & 3u, address = (t.sampler >> 2) & 3u;
    const float sx = (float)t.width, sy = (float)t.height;
    if (filter == 0u) return lc_texture2d_read<lc_float4>(t, lc_uint2((uint32_t)lc_sample_coord(address, uv.x, sx), (uint32_t)lc_sample_coord(address, uv.y, sy)));
    const float ax = lc_sample_coord(address, uv.x - 0.5f * (1.0f / sx), sx), bx = lc_sample_coord(address, uv.x + 0.5f * (1.0f / sx), sx);
    const float ay = lc_sample_coord(address, uv.y - 0.5f * (1.0f / sy), sy), by = lc_sample_coord(address, uv.y + 0.5f * (1.0f / sy), sy);
    const float x0 = fminf(ax, bx), x1 = fmaxf(ax, bx), y0 = fminf(ay, by), y1 = fmaxf(ay, by);
    const float tx = lc_fract(x1), ty = lc_fract(y1);
    const lc_uint2 c0((uint32_t)x0, (uint32_t)y0), c1((uint32_t)x1, (uint32_t)y1);
    const lc_float4 v00 = lc_texture2d_read<lc_float4>(t, c0), v01 = lc_texture2d_read<lc_float4>(t, lc_uint2(c1.x, c0.y));
    const lc_float4 v10 = lc_texture2d_read<lc_float4>(t, lc_uint2(c0.x, c1.y)), v11 = lc_texture2d_read<lc_float4>(t, c1);
    return lc_lerp(lc_lerp(v00, v01, tx), lc_lerp(v10, v11, tx), ty);
}
__device__ inline lc_float4 lc_texture3d_sample(const lc_texture &t, lc_float3 uvw) {
    const uint32_t filter = t.sampler & 3u, address = (t.sampler >> 2) & 3u;
    const float s[3] = {(float)t.width, (float)t.height, (float)t.depth}, u[3] = {uvw.x, uvw.y, uvw.z};
    if (filter == 0u) return lc_texture3d_read<lc_float4>(t, lc_uint3((uint32_t)lc_sample_coord(address, u[0], s[0]), (uint32_t)lc_sample_coord(address, u[1], s[1]), (uint32_t)lc_sample_coord(address, u[2], s[2])));
    float lo[3], hi[3];
    for (int k = 0; k < 3; k++) {
        const float a = lc_sample_coord(address, u[k] - 0.5f * (1.0f / s[k]), s[k]), b = lc_sample_coord(address, u[k] + 0.5f * (1.0f / s[k]), s[k]);
        lo[k] = fminf(a, b); hi[k] = fmaxf(a, b);
    }
    const float tx = lc_fract(hi[0]), ty = lc_fract(hi[1]), tz = lc_fract(hi[2]);
    const uint32_t x0 = (uint32_t)lo[0], y0 = (uint32_t)lo[1], z0 = (uint32_t)lo[2], x1 = (uint32_t)hi[0], y1 = (uint32_t)hi[1], z1 = (uint32_t)hi[2];
    auto rd = [&](uint32_t x, uint32_t y, uint32_t z) { return lc_texture3d_read<lc_float4>(t, lc_uint3(x, y, z)); };
    return lc_lerp(lc_lerp(lc_lerp(rd(x0, y0, z0), rd(x1, y0, z0), tx), lc_lerp(rd(x0, y1, z0), rd(x1, y1, z0), tx), ty),
                   lc_lerp(lc_lerp(rd(x0, y0, z1), rd(x1, y0, z1), tx), lc_lerp(rd(x0, y1, z1), rd(x1, y1, z1), tx), ty), tz);
}
__device__ inline lc_float4 lc_bindless_texture2d_sample(const lc_bindless &a, uint32_t slot, lc_float2 uv) { return lc_texture2d_sample(a.slots[slot].tex2d, uv); }
__device__ inline lc_float4 lc_bindless_texture3d_sample(const lc_bindless &a, uint32_t slot, lc_float3 uvw) { return lc_texture3d_sample(a.slots[slot].tex3d, uvw); }
__device__ inline lc_uint3 lc_bindless_texture3d_size(const lc_bindless &a, uint32_t slot) { return lc_texture3d_size(a.slots[slot].tex3d); }

// ---- ray tracing (rows 5-8 of SURVEY.md §8a) ------------------------------------------------------------------------------
// Ray {orig:[f32;3], tmin, dir:[f32;3], tmax} 32 B (rtx.rs:329-338); hit record {inst, prim, bary:Float2, committed_ray_t}
// (rtx.rs:356-366).  The generated code bit-casts the IR struct values to and from these, as cpp.rs:1334-1352 does.
struct alignas(16) lc_ray_rec { float o[3]; float tmin; float d[3]; float tmax; };
struct alignas(8) lc_hit_rec { uint32_t inst, prim; float u, v; float t; uint32_t pad; };
__device__ inline lc_hit_rec lc_trace_closest(const lc_accel &a, const lc_ray_rec &r, uint32_t mask) {
    const lcb::DeviceHit h = lcb::trace_one<false>(a.view, make_float4(r.o[0], r.o[1], r.o[2], r.tmin), make_float4(r.d[0], r.d[1], r.d[2], r.tmax), mask);
    lc_hit_rec out;
    out.inst = h.inst; out.prim = h.prim; out.u = h.u; out.v = h.v; out.t = h.t; out.pad = 0u;
    return out;
}
__device__ inline bool lc_trace_any(const lc_accel &a, const lc_ray_rec &r, uint32_t mask) {
    return lcb::trace_one<true>(a.view, make_float4(r.o[0], r.o[1], r.o[2], r.tmin), make_float4(r.d[0], r.d[1], r.d[2], r.tmax), mask).inst != lcb::kNone;
}
// Wavefront lowering (ir_lower.cpp header): a trace call parks its ray in the lane's traversal state and yields; the result is read
// back where the body resumes.  Same records, same arithmetic and the same barycentrics as lc_trace_closest / lc_trace_any above.
__device__ inline bool lc_wave_begin(lcb::WaveLane &w, lcb::WaveShared &s, const lc_accel &a, const lc_ray_rec &r, uint32_t mask, bool any) {
    return lcb::wave_begin(w, s, a.view, make_float4(r.o[0], r.o[1], r.o[2], r.tmin), make_float4(r.d[0], r.d[1], r.d[2], r.tmax), mask, any);
}
__device__ inline bool lc_wave_any(const lcb::WaveLane &w) { return w.hit_inst != lcb::kNone; }
__device__ inline lc_hit_rec lc_wave_closest(const lcb::WaveLane &w, const lcb::WaveShared &s, const lc_accel &a) {
    const lcb::DeviceHit h = lcb::wave_closest_result(w, s, a.view);
    lc_hit_rec out;
    out.inst = h.inst; out.prim = h.prim; out.u = h.u; out.v = h.v; out.t = h.t; out.pad = 0u;
    return out;
}
// RayQuery objects (defs::RayQuery, cpu_kernel_defs/src/lib.rs:127-141; accessors cpu_resource.h:318-396).  CommittedHit is
// {inst, prim, bary, hit_type (0 miss / 1 triangle / 2 procedural), committed_ray_t} (defs:68-77); a query that commits nothing
// keeps its initial record: inst = prim = ~0, everything else zero (cpu_resource.h:322-331).
struct alignas(8) lc_committed_hit { uint32_t inst, prim; float u, v; uint32_t hit_type; float t; };
struct lc_procedural_rec { uint32_t inst, prim; };  // defs::ProceduralHit (defs:100-105)
struct lc_ray_query_state {
    const lc_accel *accel; lc_ray_rec ray; uint32_t mask; bool terminate_on_first;
    lc_hit_rec cur_triangle; lc_procedural_rec cur_procedural; float cur_committed_t; bool cur_committed, terminated;
    lc_committed_hit hit;
};
__device__ inline lc_ray_query_state lc_make_ray_query(const lc_accel &a, const lc_ray_rec &r, uint32_t mask, bool any) {
    lc_ray_query_state q;
    q.accel = &a; q.ray = r; q.mask = mask; q.terminate_on_first = any;
    q.cur_triangle = lc_hit_rec{~0u, ~0u, 0.f, 0.f, 0.f, 0u}; q.cur_procedural = lc_procedural_rec{~0u, ~0u}; q.cur_committed_t = 0.f;
    q.cur_committed = false; q.terminated = false;
    q.hit = lc_committed_hit{~0u, ~0u, 0.f, 0.f, 0u, 0.f};
    return q;
}
__device__ inline lc_ray_query_state lc_ray_query_all(const lc_accel &a, const lc_ray_rec &r, uint32_t mask) { return lc_make_ray_query(a, r, mask, false); }
__device__ inline lc_ray_query_state lc_ray_query_any(const lc_accel &a, const lc_ray_rec &r, uint32_t mask) { return lc_make_ray_query(a, r, mask, true); }
// Instruction::RayQuery: run the traversal.  Triangles of non-opaque instances go through on_triangle, AABBs of procedural
// instances through on_procedural; either may call RayQueryCommit* / RayQueryTerminate on the same object — the contract of
// AccelImpl::ray_query's filter_fn / intersect_fn (cpu/accel.rs:646-757).  Both see rq.ray with tmax = the closest committed t so far.
template <class OnTriangle, class OnProcedural> struct lc_query_hook {
    lc_ray_query_state &q; OnTriangle &on_triangle; OnProcedural &on_procedural;
    __device__ int triangle(uint32_t inst, uint32_t prim, float u, float v, float t) {
        q.cur_triangle = lc_hit_rec{inst, prim, u, v, t, 0u};
        q.ray.tmax = t;  // accel.rs:668: the filter sees the candidate's t as tfar
        q.cur_committed = false; q.terminated = false;
        on_triangle();
        return (q.cur_committed ? 1 : 0) | (q.terminated ? 2 : 0);
    }
    __device__ int procedural(uint32_t inst, uint32_t prim, float t_far, float &t) {
        q.cur_procedural = lc_procedural_rec{inst, prim};
        q.ray.tmax = t_far;  // accel.rs:738
        q.cur_committed = false; q.terminated = false;
        on_procedural();
        t = q.cur_committed_t;
        return (q.cur_committed ? 1 : 0) | (q.terminated ? 2 : 0);
    }
};
template <class OnTriangle, class OnProcedural> __device__ inline void lc_ray_query(lc_ray_query_state &q, OnTriangle on_triangle, OnProcedural on_procedural) {
    lc_query_hook<OnTriangle, OnProcedural> hook{q, on_triangle, on_procedural};
    const float4 ra = make_float4(q.ray.o[0], q.ray.o[1], q.ray.o[2], q.ray.tmin), rb = make_float4(q.ray.d[0], q.ray.d[1], q.ray.d[2], q.ray.tmax);
    const lcb::DeviceHit h = lcb::trace_one_impl<false, true>(q.accel->view, ra, rb, q.mask, q.terminate_on_first, hook);
    if (h.inst != lcb::kNone) q.hit = lc_committed_hit{h.inst, h.prim, h.u, h.v, h.kind, h.t};
}

// instance accessors: the transform is returned as the column-major Mat4 built from the row-major 3x4 (stream.rs:582-595)
__device__ inline lc_float4x4 lc_accel_instance_transform(const lc_accel &a, uint32_t i) {
    const float *m = a.view.instances[i].affine;
    return lc_make_mat(lc_float4(m[0], m[4], m[8], 0.f), lc_float4(m[1], m[5], m[9], 0.f), lc_float4(m[2], m[6], m[10], 0.f), lc_float4(m[3], m[7], m[11], 1.f));
}
__device__ inline uint32_t lc_accel_instance_visibility_mask(const lc_accel &a, uint32_t i) { return a.view.instances[i].visibility; }
__device__ inline uint32_t lc_accel_instance_user_id(const lc_accel &a, uint32_t i) { return a.view.instances[i].user_id; }
// Setters (cpu/accel.rs:560-579, stream.rs:596-658) edit the device's instance table and raise the accel's dirty flag: the next
// AccelBuild reads the edited slots back into the host mirror before it rebuilds the TLAS.  Visibility, opacity and user id take
// effect for traversals at once; a new transform moves the instance's box only at that next AccelBuild, as on the CPU backend.
__device__ inline void lc_set_instance_visibility(const lc_accel &a, uint32_t i, uint32_t m) { a.instances_rw[i].visibility = m; *a.dirty = 1u; }
__device__ inline void lc_set_instance_user_id(const lc_accel &a, uint32_t i, uint32_t id) { a.instances_rw[i].user_id = id; *a.dirty = 1u; }
__device__ inline void lc_set_instance_opacity(const lc_accel &a, uint32_t i, bool opaque) {
    const uint32_t f = a.instances_rw[i].flags;
    a.instances_rw[i].flags = opaque ? (f | 2u) : (f & ~2u);
    *a.dirty = 1u;
}
__device__ inline void lc_set_instance_transform(const lc_accel &a, uint32_t i, const lc_float4x4 &m) {
    lcb::InstanceRec &r = a.instances_rw[i];
    float aff[12];
    for (int row = 0; row < 3; row++) for (int col = 0; col < 4; col++) aff[4 * row + col] = m[col][row];  // column-major Mat4 -> row-major 3x4
    for (int k = 0; k < 12; k++) r.affine[k] = aff[k];
    // world -> object: double-precision adjugate inverse rounded once to fp32, the same formula the host uses (device.cu invert_affine)
    const double A = aff[0], B = aff[1], Cc = aff[2], D = aff[4], E = aff[5], F = aff[6], G = aff[8], H = aff[9], I = aff[10];
    const double tx = aff[3], ty = aff[7], tz = aff[11];
    const double c00 = E * I - F * H, c01 = Cc * H - B * I, c02 = B * F - Cc * E;
    const double c10 = F * G - D * I, c11 = A * I - Cc * G, c12 = Cc * D - A * F;
    const double c20 = D * H - E * G, c21 = B * G - A * H, c22 = A * E - B * D;
    const double rdet = 1.0 / (A * c00 + B * c10 + Cc * c20);
    const double n00 = c00 * rdet, n01 = c01 * rdet, n02 = c02 * rdet, n10 = c10 * rdet, n11 = c11 * rdet, n12 = c12 * rdet, n20 = c20 * rdet, n21 = c21 * rdet, n22 = c22 * rdet;
    r.inv[0] = (float)n00; r.inv[1] = (float)n01; r.inv[2] = (float)n02; r.inv[3] = (float)(-(n00 * tx + n01 * ty + n02 * tz));
    r.inv[4] = (float)n10; r.inv[5] = (float)n11; r.inv[6] = (float)n12; r.inv[7] = (float)(-(n10 * tx + n11 * ty + n12 * tz));
    r.inv[8] = (float)n20; r.inv[9] = (float)n21; r.inv[10] = (float)n22; r.inv[11] = (float)(-(n20 * tx + n21 * ty + n22 * tz));
    bool identity = true;  // flag bit 4: exactly the identity, zeros of either sign (the host does the same in AccelBuild)
    for (int k = 0; k < 12; k++) identity = identity && r.inv[k] == ((k == 0 || k == 5 || k == 10) ? 1.0f : 0.0f);
    r.flags = identity ? (r.flags | lcb::kInstIdentity) : (r.flags & ~lcb::kInstIdentity);
    *a.dirty = 1u;
}
