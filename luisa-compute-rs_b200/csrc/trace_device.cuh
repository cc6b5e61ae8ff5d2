// trace_device.cuh — the per-ray device pieces of the traversal, shared by the batch kernels (trace.cu) and by kernels that
// trace from inside their own threads (path_tracer.cu; an IR-lowered DSL kernel would include this header the same way
// the CPU backend's generated code includes cpu_resource.h and calls lc_trace_closest / lc_trace_any, cpu_resource.h:288-294).
#pragma once
#ifdef __CUDACC_RTC__
#include "rt_types.cuh"
#else
#include "build.cuh"
#endif

namespace lcb {

constexpr uint32_t kNone = 0xffffffffu;

struct RaySetup {
    float ox, oy, oz, dx, dy, dz;  // ray in the current space (world or object)
    float ix, iy, iz;              // clamped reciprocal direction for slab tests
    float sx, sy, sz;              // shear constants of the canonical triangle test
    int kz;
    uint32_t octinv;               // 7 ^ sign bits (bit k = direction negative along k)
};

__device__ __forceinline__ float safe_rcp(float d) {
    // |d| < 1e-20 is treated as 1e-20 with d's sign bit: keeps slab arithmetic finite; culling stays conservative
    float a = fabsf(d) < 1e-20f ? copysignf(1e-20f, d) : d;
    return __frcp_rn(a);
}

__device__ __forceinline__ void finish_setup(RaySetup &r) {
    r.ix = safe_rcp(r.dx); r.iy = safe_rcp(r.dy); r.iz = safe_rcp(r.dz);
    const uint32_t sgn = (__float_as_uint(r.dx) >> 31) | ((__float_as_uint(r.dy) >> 31) << 1) | ((__float_as_uint(r.dz) >> 31) << 2);
    r.octinv = 7u ^ sgn;
    int kz = 0;
    float m = fabsf(r.dx);
    if (fabsf(r.dy) > m) { kz = 1; m = fabsf(r.dy); }
    if (fabsf(r.dz) > m) { kz = 2; }
    r.kz = kz;
    const float dz = kz == 0 ? r.dx : (kz == 1 ? r.dy : r.dz);
    const float dx = kz == 0 ? r.dy : (kz == 1 ? r.dz : r.dx);
    const float dy = kz == 0 ? r.dz : (kz == 1 ? r.dx : r.dy);
    r.sx = __fdiv_rn(dx, dz);
    r.sy = __fdiv_rn(dy, dz);
    r.sz = __frcp_rn(dz);
}

__device__ __forceinline__ void setup_world(RaySetup &r, const float4 a, const float4 b) {
    r.ox = a.x; r.oy = a.y; r.oz = a.z; r.dx = b.x; r.dy = b.y; r.dz = b.z;
    finish_setup(r);
}

// world -> object with the canonical nested-fma order
__device__ __forceinline__ void transform_ray(RaySetup &r, const float4 wo, const float4 wd, const float4 m0, const float4 m1, const float4 m2) {
    r.ox = __fmaf_rn(m0.x, wo.x, __fmaf_rn(m0.y, wo.y, __fmaf_rn(m0.z, wo.z, m0.w)));
    r.oy = __fmaf_rn(m1.x, wo.x, __fmaf_rn(m1.y, wo.y, __fmaf_rn(m1.z, wo.z, m1.w)));
    r.oz = __fmaf_rn(m2.x, wo.x, __fmaf_rn(m2.y, wo.y, __fmaf_rn(m2.z, wo.z, m2.w)));
    r.dx = __fmaf_rn(m0.x, wd.x, __fmaf_rn(m0.y, wd.y, __fmul_rn(m0.z, wd.z)));
    r.dy = __fmaf_rn(m1.x, wd.x, __fmaf_rn(m1.y, wd.y, __fmul_rn(m1.z, wd.z)));
    r.dz = __fmaf_rn(m2.x, wd.x, __fmaf_rn(m2.y, wd.y, __fmul_rn(m2.z, wd.z)));
}

__device__ __forceinline__ void setup_object(RaySetup &r, const float4 wo, const float4 wd, const float4 m0, const float4 m1, const float4 m2) {
    transform_ray(r, wo, wd, m0, m1, m2);
    finish_setup(r);
}

// Entering an instance from world space, where `r` holds the world-space setup.  Instances whose world -> object matrix is exactly the
// identity (zeros of either sign) carry flag bit 4 (set by the host in AccelBuild, kept by lc_set_instance_transform): for them the canonical transform
// fma(1, x, fma(0, y, fma(0, z, 0))) reproduces every component bit for bit as long as none of the six is zero or non-finite (a zero
// could change sign, 0 * inf is NaN), so the object-space setup — a pure function of those bits — is the world-space one and nothing is
// recomputed: no matrix load, no divisions.  Single-instance scenes and the Cornell box's eight meshes take this path for practically
// every ray.  Returns true in that case; the caller records it in the stack sentinel so that leaving the instance skips the
// world-space re-setup too.
constexpr uint32_t kInstIdentity = 16u;
__device__ __forceinline__ bool nonzero_finite(float x) { return ((__float_as_uint(x) & 0x7fffffffu) - 1u) < 0x7f7fffffu; }
__device__ __forceinline__ bool enter_instance(RaySetup &r, uint32_t inst_flags, const float4 *__restrict__ rec) {
#ifndef LCB_NO_IDENTITY_SHORTCUT  // A/B switch for kernel sweeps
    if ((inst_flags & kInstIdentity) && nonzero_finite(r.ox) && nonzero_finite(r.oy) && nonzero_finite(r.oz) && nonzero_finite(r.dx) && nonzero_finite(r.dy) &&
        nonzero_finite(r.dz))
        return true;
#endif
    const float4 m0 = __ldg(rec), m1 = __ldg(rec + 1), m2 = __ldg(rec + 2);
    setup_object(r, make_float4(r.ox, r.oy, r.oz, 0.f), make_float4(r.dx, r.dy, r.dz, 0.f), m0, m1, m2);
    return false;
}

// 16-bit plane index -> float 2^23 + q in ONE byte-permute (no int->float conversion, no subtraction): the 2^23
// bias is folded into the per-node plane offsets below, at the price of half a quantisation step of rounding
// slop that the padding absorbs (one extra step, 2^-16 of the node extent).  The constant sits in the first
// operand so that the selector is an immediate and 0x4B000000 lives in one register for the whole kernel.
__device__ __forceinline__ float q16_lo(uint32_t w) { return __uint_as_float(__byte_perm(0x4B000000u, w, 0x3254)); }
__device__ __forceinline__ float q16_hi(uint32_t w) { return __uint_as_float(__byte_perm(0x4B000000u, w, 0x3276)); }

// 256-bit read-only global load (LDG.E.ENL2.256.CONSTANT): one full 32-byte sector per lane and instruction
struct U8 { uint32_t v[8]; };
__device__ __forceinline__ U8 ldg256(const void *p) {
    U8 r;
    asm("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7])
                 : "l"(p));
    return r;
}

// each byte -> 0xff if its top bit is set, else 0x00 (prmt sign-replicate mode; __byte_perm masks the selector's msb away)
__device__ __forceinline__ uint32_t sign_extend_bytes(uint32_t x) {
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(x), "r"(0u), "r"(0xba98u));
    return d;
}

// 8-bit plane index (byte B of word W) -> float 1 + q * 2^-15 in ONE byte-permute: the byte lands in bits 8..15 of the mantissa of 1.0f.
// Folding the 1 into the per-node constants then costs an ulp of 2^15 * step = 2^-9 of a quantisation step (the 16-bit form above pays
// half a step, because 2^23 + q leaves no mantissa below q), so the slop the padding has to absorb shrinks from a whole step to 1/64.
#define LCB_Q8(W, B) __uint_as_float(__byte_perm(0x3F800000u, W, 0x3240 + ((B) << 4)))

// Tests the 8 children of one node.  Returns the hit mask: bits 24..31 internal children in
// traversal priority order for this ray's octant, bits 0..23 leaf primitives.
__device__ __forceinline__ uint32_t intersect_node(const WideNode *__restrict__ node, const RaySetup &r, float tmin, float tmax,
                                                   uint32_t &child_base, uint32_t &prim_base, uint32_t &imask) {
    const uint4 *p = reinterpret_cast<const uint4 *>(node);
    const bool neg_x = (r.octinv & 1u) == 0, neg_y = (r.octinv & 2u) == 0, neg_z = (r.octinv & 4u) == 0;
#if LCB_NODE_BITS == 16
    const U8 H = ldg256(p), X = ldg256(p + 2), Y = ldg256(p + 4), Z = ldg256(p + 6);
    // near/far plane vectors by direction sign (words 0..3 = lower planes, 4..7 = upper planes of the 8 slots)
#define LCB_NEAR(V, NEG) make_uint4(NEG ? V.v[4] : V.v[0], NEG ? V.v[5] : V.v[1], NEG ? V.v[6] : V.v[2], NEG ? V.v[7] : V.v[3])
#define LCB_FAR(V, NEG) make_uint4(NEG ? V.v[0] : V.v[4], NEG ? V.v[1] : V.v[5], NEG ? V.v[2] : V.v[6], NEG ? V.v[3] : V.v[7])
    const uint4 qnx = LCB_NEAR(X, neg_x), qfx = LCB_FAR(X, neg_x);
    const uint4 qny = LCB_NEAR(Y, neg_y), qfy = LCB_FAR(Y, neg_y);
    const uint4 qnz = LCB_NEAR(Z, neg_z), qfz = LCB_FAR(Z, neg_z);
#undef LCB_NEAR
#undef LCB_FAR
#else
    // three sectors: header | x lo, x hi, y lo, y hi (8 bytes each) | z lo, z hi, spare
    const U8 H = ldg256(p), XY = ldg256(p + 2);
    const uint4 Zq = __ldg(p + 4);
    const uint2 qnx = neg_x ? make_uint2(XY.v[2], XY.v[3]) : make_uint2(XY.v[0], XY.v[1]), qfx = neg_x ? make_uint2(XY.v[0], XY.v[1]) : make_uint2(XY.v[2], XY.v[3]);
    const uint2 qny = neg_y ? make_uint2(XY.v[6], XY.v[7]) : make_uint2(XY.v[4], XY.v[5]), qfy = neg_y ? make_uint2(XY.v[4], XY.v[5]) : make_uint2(XY.v[6], XY.v[7]);
    const uint2 qnz = neg_z ? make_uint2(Zq.z, Zq.w) : make_uint2(Zq.x, Zq.y), qfz = neg_z ? make_uint2(Zq.x, Zq.y) : make_uint2(Zq.z, Zq.w);
#endif
    const uint4 n0 = make_uint4(H.v[0], H.v[1], H.v[2], H.v[3]), n1 = make_uint4(H.v[4], H.v[5], H.v[6], H.v[7]);
    child_base = n1.x; prim_base = n1.y; imask = n0.w >> 24;
    const float sclx = __uint_as_float((n0.w & 0xffu) << 23), scly = __uint_as_float((n0.w & 0xff00u) << 15), sclz = __uint_as_float((n0.w & 0xff0000u) << 7);
    const float rx = __uint_as_float(n0.x) - r.ox, ry = __uint_as_float(n0.y) - r.oy, rz = __uint_as_float(n0.z) - r.oz;
    // conservative padding: 2^-20 of the L-inf distance from the ray origin to the far side of the node frame
    constexpr float kSpan = (float)kQMax;
    const float R = fmaxf(fmaxf(fabsf(rx) + kSpan * sclx, fabsf(ry) + kSpan * scly), fabsf(rz) + kSpan * sclz);
    const float pad = R * (1.0f / 1048576.0f);
    const float ax = sclx * r.ix, ay = scly * r.iy, az = sclz * r.iz;
    const float cx = rx * r.ix, cy = ry * r.iy, cz = rz * r.iz;
#if LCB_NODE_BITS == 16
    const float px = fmaf(pad, fabsf(r.ix), fabsf(ax)), py = fmaf(pad, fabsf(r.iy), fabsf(ay)), pz = fmaf(pad, fabsf(r.iz), fabsf(az));
    const float bnx = fmaf(-8388608.0f, ax, cx - px), bfx = fmaf(-8388608.0f, ax, cx + px);
    const float bny = fmaf(-8388608.0f, ay, cy - py), bfy = fmaf(-8388608.0f, ay, cy + py);
    const float bnz = fmaf(-8388608.0f, az, cz - pz), bfz = fmaf(-8388608.0f, az, cz + pz);
    const float mx = ax, my = ay, mz = az;
#else
    const float px = fmaf(pad, fabsf(r.ix), fabsf(ax) * (1.0f / 64.0f)), py = fmaf(pad, fabsf(r.iy), fabsf(ay) * (1.0f / 64.0f)), pz = fmaf(pad, fabsf(r.iz), fabsf(az) * (1.0f / 64.0f));
    const float mx = ax * 32768.0f, my = ay * 32768.0f, mz = az * 32768.0f;  // plane index arrives as 1 + q * 2^-15
    const float bnx = (cx - px) - mx, bfx = (cx + px) - mx;
    const float bny = (cy - py) - my, bfy = (cy + py) - my;
    const float bnz = (cz - pz) - mz, bfz = (cz + pz) - mz;
#endif
    // hit-mask construction on 4 meta bytes at a time (after Ylitie et al. 2017): per child only a byte extract,
    // a shift and a select remain.  Internal children (low 5 bits >= 24) get their bit index XORed with the
    // ray octant so that __clz order is front-to-back order.
    const uint32_t oct4 = r.octinv * 0x01010101u;
    const uint32_t inner_lo = sign_extend_bytes((n1.z & (n1.z << 1) & 0x10101010u) << 3);  // 0xff per internal child
    const uint32_t inner_hi = sign_extend_bytes((n1.w & (n1.w << 1) & 0x10101010u) << 3);
    const uint32_t idx_lo = (n1.z ^ (oct4 & inner_lo)) & 0x1f1f1f1fu, idx_hi = (n1.w ^ (oct4 & inner_hi)) & 0x1f1f1f1fu;
    const uint32_t bits_lo = (n1.z >> 5) & 0x07070707u, bits_hi = (n1.w >> 5) & 0x07070707u;
    uint32_t hits = 0;
#define LCB_CHILD_T(NX, NY, NZ, FX, FY, FZ, BITS, IDX, SHIFT)                                                \
    {                                                                                                        \
        const float tn = fmaxf(fmaxf(fmaf(NX, mx, bnx), fmaf(NY, my, bny)), fmaxf(fmaf(NZ, mz, bnz), tmin));   \
        const float tf = fminf(fminf(fmaf(FX, mx, bfx), fmaf(FY, my, bfy)), fminf(fmaf(FZ, mz, bfz), tmax));   \
        const uint32_t b = ((BITS >> SHIFT) & 0xffu) << ((IDX >> SHIFT) & 31u);                              \
        hits |= tn <= tf ? b : 0u;                                                                           \
    }
#if LCB_NODE_BITS == 16
#define LCB_CHILD(WORD, CONV, BITS, IDX, SHIFT) LCB_CHILD_T(CONV(qnx.WORD), CONV(qny.WORD), CONV(qnz.WORD), CONV(qfx.WORD), CONV(qfy.WORD), CONV(qfz.WORD), BITS, IDX, SHIFT)
    LCB_CHILD(x, q16_lo, bits_lo, idx_lo, 0)
    LCB_CHILD(x, q16_hi, bits_lo, idx_lo, 8)
    LCB_CHILD(y, q16_lo, bits_lo, idx_lo, 16)
    LCB_CHILD(y, q16_hi, bits_lo, idx_lo, 24)
    LCB_CHILD(z, q16_lo, bits_hi, idx_hi, 0)
    LCB_CHILD(z, q16_hi, bits_hi, idx_hi, 8)
    LCB_CHILD(w, q16_lo, bits_hi, idx_hi, 16)
    LCB_CHILD(w, q16_hi, bits_hi, idx_hi, 24)
#else
#define LCB_CHILD(WORD, B, BITS, IDX, SHIFT) LCB_CHILD_T(LCB_Q8(qnx.WORD, B), LCB_Q8(qny.WORD, B), LCB_Q8(qnz.WORD, B), LCB_Q8(qfx.WORD, B), LCB_Q8(qfy.WORD, B), LCB_Q8(qfz.WORD, B), BITS, IDX, SHIFT)
    LCB_CHILD(x, 0, bits_lo, idx_lo, 0)
    LCB_CHILD(x, 1, bits_lo, idx_lo, 8)
    LCB_CHILD(x, 2, bits_lo, idx_lo, 16)
    LCB_CHILD(x, 3, bits_lo, idx_lo, 24)
    LCB_CHILD(y, 0, bits_hi, idx_hi, 0)
    LCB_CHILD(y, 1, bits_hi, idx_hi, 8)
    LCB_CHILD(y, 2, bits_hi, idx_hi, 16)
    LCB_CHILD(y, 3, bits_hi, idx_hi, 24)
#endif
#undef LCB_CHILD
#undef LCB_CHILD_T
    return hits;
}

__device__ __forceinline__ float pick(int k, float x, float y, float z) { return k == 0 ? x : (k == 1 ? y : z); }

// The canonical fp32 ray/triangle evaluation (oracle.c canon_tri).  The traversal loop only needs
// the decision and t; (V, W, det) are handed back so that the barycentrics can be formed where they are needed.
__device__ __forceinline__ bool canonical_triangle(const RaySetup &r, float tmin, float tmax, const float4 v0, const float4 v1, const float4 v2,
                                                   float &t_out, float &V_out, float &W_out, float &det_out) {
    const float a0 = __fsub_rn(v0.x, r.ox), a1 = __fsub_rn(v0.y, r.oy), a2 = __fsub_rn(v0.z, r.oz);
    const float b0 = __fsub_rn(v1.x, r.ox), b1 = __fsub_rn(v1.y, r.oy), b2 = __fsub_rn(v1.z, r.oz);
    const float c0 = __fsub_rn(v2.x, r.ox), c1 = __fsub_rn(v2.y, r.oy), c2 = __fsub_rn(v2.z, r.oz);
    const int kz = r.kz;
    const float a_z = pick(kz, a0, a1, a2), a_x = pick(kz, a1, a2, a0), a_y = pick(kz, a2, a0, a1);
    const float b_z = pick(kz, b0, b1, b2), b_x = pick(kz, b1, b2, b0), b_y = pick(kz, b2, b0, b1);
    const float c_z = pick(kz, c0, c1, c2), c_x = pick(kz, c1, c2, c0), c_y = pick(kz, c2, c0, c1);
    const float ax = __fmaf_rn(-r.sx, a_z, a_x), ay = __fmaf_rn(-r.sy, a_z, a_y);
    const float bx = __fmaf_rn(-r.sx, b_z, b_x), by = __fmaf_rn(-r.sy, b_z, b_y);
    const float cx = __fmaf_rn(-r.sx, c_z, c_x), cy = __fmaf_rn(-r.sy, c_z, c_y);
    float U = __fsub_rn(__fmul_rn(cx, by), __fmul_rn(cy, bx));
    float V = __fsub_rn(__fmul_rn(ax, cy), __fmul_rn(ay, cx));
    float W = __fsub_rn(__fmul_rn(bx, ay), __fmul_rn(by, ax));
    if (U == 0.0f || V == 0.0f || W == 0.0f) {
        U = __double2float_rn(__dsub_rn(__dmul_rn((double)cx, (double)by), __dmul_rn((double)cy, (double)bx)));
        V = __double2float_rn(__dsub_rn(__dmul_rn((double)ax, (double)cy), __dmul_rn((double)ay, (double)cx)));
        W = __double2float_rn(__dsub_rn(__dmul_rn((double)bx, (double)ay), __dmul_rn((double)by, (double)ax)));
    }
    if ((U < 0.0f || V < 0.0f || W < 0.0f) && (U > 0.0f || V > 0.0f || W > 0.0f)) return false;
    const float det = __fadd_rn(__fadd_rn(U, V), W);
    if (det == 0.0f) return false;
    const float az = __fmul_rn(r.sz, a_z), bz = __fmul_rn(r.sz, b_z), cz = __fmul_rn(r.sz, c_z);
    const float T = __fmaf_rn(U, az, __fmaf_rn(V, bz, __fmul_rn(W, cz)));
    const float t = __fdiv_rn(T, det);
    if (!(t > tmin && t <= tmax)) return false;
    t_out = t; V_out = V; W_out = W; det_out = det;
    return true;
}

// Reported barycentrics of the winning triangle: one double-precision Moeller-Trumbore evaluation on the
// canonical object-space ray (fixed operation order; oracle.c refine_bary).  Once per ray, off the hot loop.
// Returns false when the double determinant vanishes (the canonical fp32 barycentrics stand).
__device__ __forceinline__ bool refine_bary(const RaySetup &r, const float4 v0, const float4 v1, const float4 v2, float &u_out, float &v_out) {
    const double e1x = __dsub_rn((double)v1.x, (double)v0.x), e1y = __dsub_rn((double)v1.y, (double)v0.y), e1z = __dsub_rn((double)v1.z, (double)v0.z);
    const double e2x = __dsub_rn((double)v2.x, (double)v0.x), e2y = __dsub_rn((double)v2.y, (double)v0.y), e2z = __dsub_rn((double)v2.z, (double)v0.z);
    const double sx = __dsub_rn((double)r.ox, (double)v0.x), sy = __dsub_rn((double)r.oy, (double)v0.y), sz = __dsub_rn((double)r.oz, (double)v0.z);
    const double dx = r.dx, dy = r.dy, dz = r.dz;
    const double px = __dsub_rn(__dmul_rn(dy, e2z), __dmul_rn(dz, e2y));
    const double py = __dsub_rn(__dmul_rn(dz, e2x), __dmul_rn(dx, e2z));
    const double pz = __dsub_rn(__dmul_rn(dx, e2y), __dmul_rn(dy, e2x));
    const double det = __dadd_rn(__dadd_rn(__dmul_rn(e1x, px), __dmul_rn(e1y, py)), __dmul_rn(e1z, pz));
    if (det == 0.0) return false;
    const double inv = __ddiv_rn(1.0, det);
    const double u = __dmul_rn(__dadd_rn(__dadd_rn(__dmul_rn(sx, px), __dmul_rn(sy, py)), __dmul_rn(sz, pz)), inv);
    const double qx = __dsub_rn(__dmul_rn(sy, e1z), __dmul_rn(sz, e1y));
    const double qy = __dsub_rn(__dmul_rn(sz, e1x), __dmul_rn(sx, e1z));
    const double qz = __dsub_rn(__dmul_rn(sx, e1y), __dmul_rn(sy, e1x));
    const double v = __dmul_rn(__dadd_rn(__dadd_rn(__dmul_rn(dx, qx), __dmul_rn(dy, qy)), __dmul_rn(dz, qz)), inv);
    u_out = __double2float_rn(u); v_out = __double2float_rn(v);
    return true;
}

// ---- curves (CurveBuild, api_types:620-631; GeometryImpl::build_curve, cpu/accel.rs:142-203) -------------------------------
// Control points are float4 (x, y, z, radius); segment i starts at control point seg[i] and uses 2 (PiecewiseLinear) or 4 of them.
// The surface is the sweep of a sphere of radius r(u) along c(u) — the union of those spheres.  A segment is put into the power
// basis with the frontend's own matrices (CubicCurve::{bspline, catmull_rom, bezier}, lc/src/rtx/curve.rs:88-139, so `u` means what
// CurveEvaluator expects), cubic segments are cut at u = k/8 into kCurveSubdiv pieces, and every piece is a rounded cone between
// the spheres at its ends — one BLAS leaf (CurveSeg below, in the PackedTri slot) with its own box.  The reference's round curves
// (Embree RTC_GEOMETRY_TYPE_ROUND_*_CURVE) are intersected iteratively to a tolerance instead; parity unpinned, as for triangles.
// A hit reports prim = segment index, bary = (u, -1) (cpu/accel.rs:491-494) and the entry t.  oracle.c: curve_basis / canon_cone.
constexpr uint32_t kCurveLinear = 0, kCurveBSpline = 1, kCurveCatmullRom = 2, kCurveBezier = 3;  // CurveBasis, api_types:196-202

struct alignas(64) CurveSeg {   // one piece, in the 64-byte PackedTri slot
    float pa[3]; uint32_t prim;  // sphere at the piece's start; prim = segment index
    float pb[3]; float ra;       // sphere at its end; radius at the start (both radii inflated by the piece's sag bound, curve_piece)
    float rb, u0, du; uint32_t coef;  // coef: slot (in the same 64-byte array) of the segment's power-basis coefficients a0..a3, cubic bases only
    uint32_t spare[4];
};
static_assert(sizeof(CurveSeg) == 64, "CurveSeg shares the PackedTri slot");

__device__ __forceinline__ float4 curve_lin4(const float m0, const float m1, const float m2, const float m3, const float div,
                                             const float4 q0, const float4 q1, const float4 q2, const float4 q3) {
#define LCB_L4(C) __fdiv_rn(__fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(m0, q0.C), __fmul_rn(m1, q1.C)), __fmul_rn(m2, q2.C)), __fmul_rn(m3, q3.C)), div)
    return make_float4(LCB_L4(x), LCB_L4(y), LCB_L4(z), LCB_L4(w));
#undef LCB_L4
}
// power basis c(u) = ((a[0] u + a[1]) u + a[2]) u + a[3]
__device__ __forceinline__ void curve_power_basis(uint32_t basis, const float4 q0, const float4 q1, const float4 q2, const float4 q3, float4 a[4]) {
    if (basis == kCurveBSpline) {
        a[0] = curve_lin4(-1.f, 3.f, -3.f, 1.f, 6.f, q0, q1, q2, q3); a[1] = curve_lin4(3.f, -6.f, 3.f, 0.f, 6.f, q0, q1, q2, q3);
        a[2] = curve_lin4(-3.f, 0.f, 3.f, 0.f, 6.f, q0, q1, q2, q3);  a[3] = curve_lin4(1.f, 4.f, 1.f, 0.f, 6.f, q0, q1, q2, q3);
    } else if (basis == kCurveCatmullRom) {
        a[0] = curve_lin4(-1.f, 3.f, -3.f, 1.f, 2.f, q0, q1, q2, q3); a[1] = curve_lin4(2.f, -5.f, 4.f, -1.f, 2.f, q0, q1, q2, q3);
        a[2] = curve_lin4(-1.f, 0.f, 1.f, 0.f, 2.f, q0, q1, q2, q3);  a[3] = curve_lin4(0.f, 2.f, 0.f, 0.f, 2.f, q0, q1, q2, q3);
    } else {
        a[0] = curve_lin4(-1.f, 3.f, -3.f, 1.f, 1.f, q0, q1, q2, q3); a[1] = curve_lin4(3.f, -6.f, 3.f, 0.f, 1.f, q0, q1, q2, q3);
        a[2] = curve_lin4(-3.f, 3.f, 0.f, 0.f, 1.f, q0, q1, q2, q3);  a[3] = curve_lin4(1.f, 0.f, 0.f, 0.f, 1.f, q0, q1, q2, q3);
    }
}
__device__ __forceinline__ float4 curve_point(const float4 a[4], float u) {
#define LCB_H(C) __fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(a[0].C, u), a[1].C), u), a[2].C), u), a[3].C)
    return make_float4(LCB_H(x), LCB_H(y), LCB_H(z), LCB_H(w));
#undef LCB_H
}
// the two end spheres (xyz, radius) of piece k of segment `seg`
__device__ __forceinline__ void curve_piece(const uint8_t *cps, size_t cp_stride, const uint32_t *segs, uint32_t basis, uint32_t seg, uint32_t k,
                                            float4 &A, float4 &B) {
    const uint32_t first = segs[seg];
    const float4 q0 = *reinterpret_cast<const float4 *>(cps + (size_t)first * cp_stride);
    const float4 q1 = *reinterpret_cast<const float4 *>(cps + (size_t)(first + 1) * cp_stride);
    if (basis == kCurveLinear) { A = q0; B = q1; return; }
    const float4 q2 = *reinterpret_cast<const float4 *>(cps + (size_t)(first + 2) * cp_stride);
    const float4 q3 = *reinterpret_cast<const float4 *>(cps + (size_t)(first + 3) * cp_stride);
    float4 a[4];
    curve_power_basis(basis, q0, q1, q2, q3, a);
    const float u0 = (float)k * (1.0f / kCurveSubdiv), u1 = (float)(k + 1) * (1.0f / kCurveSubdiv);
    A = curve_point(a, u0);
    B = curve_point(a, u1);
    // The cone only locates the hit (refine_curve_hit decides): its radii are inflated by a bound on how far the cubic leaves its chord
    // over the piece — du^2 / 8 * max |c''| per component, c'' linear in u — so that every ray that enters the true sweep here also enters
    // the cone and becomes a candidate (oracle.c curve_piece, same operations).
    float sag = 0.0f;
#define LCB_SAG(C) sag = __fadd_rn(sag, fmaxf(fabsf(__fadd_rn(__fmul_rn(__fmul_rn(6.0f, a[0].C), u0), __fmul_rn(2.0f, a[1].C))), fabsf(__fadd_rn(__fmul_rn(__fmul_rn(6.0f, a[0].C), u1), __fmul_rn(2.0f, a[1].C)))));
    LCB_SAG(x) LCB_SAG(y) LCB_SAG(z) LCB_SAG(w)
#undef LCB_SAG
    sag = __fmul_rn(__fmul_rn(sag, 1.0f / (8.0f * kCurveSubdiv * kCurveSubdiv)), 1.0625f);
    A.w = __fadd_rn(fabsf(A.w), sag); B.w = __fadd_rn(fabsf(B.w), sag);
}

// The segment's power-basis coefficients (what refine_curve_hit evaluates), written behind the leaves of a cubic curve BLAS.
__device__ __forceinline__ void curve_coefficients(const uint8_t *cps, size_t cp_stride, const uint32_t *segs, uint32_t basis, uint32_t seg, float4 a[4]) {
    const uint32_t first = segs[seg];
    curve_power_basis(basis, *reinterpret_cast<const float4 *>(cps + (size_t)first * cp_stride), *reinterpret_cast<const float4 *>(cps + (size_t)(first + 1) * cp_stride),
                      *reinterpret_cast<const float4 *>(cps + (size_t)(first + 2) * cp_stride), *reinterpret_cast<const float4 *>(cps + (size_t)(first + 3) * cp_stride), a);
}

// Refinement of a cone hit against the true swept surface: Newton on F1 = |p - c(u)|^2 - r(u)^2 = 0, F2 = (p - c(u)).c'(u) + r(u) r'(u) = 0
// (the envelope condition), p = o + t d, in double precision with a fixed operation order — oracle.c refine_curve_hit is the same
// sequence, so both sides report the same bits.  end: -1 the piece starts the segment, +1 it ends it.  Candidates: the envelope point
// the iteration settled on (inside the segment, entry side) and, on a first / last piece, the sphere that closes the segment (exact);
// the earlier one is the hit.  Neither: the (inflated) cone only bulged out of the sweep, or the ray grazes — the piece reports nothing.
__device__ __forceinline__ bool curve_end_sphere(const double o[3], const double d[3], const float4 a[4], double ue, float &t_out, float &u_out) {
#define LCB_CE(C) __dadd_rn(__dmul_rn(__dadd_rn(__dmul_rn(__dadd_rn(__dmul_rn((double)a[0].C, ue), (double)a[1].C), ue), (double)a[2].C), ue), (double)a[3].C)
    const double c0 = LCB_CE(x), c1 = LCB_CE(y), c2 = LCB_CE(z), c3 = LCB_CE(w);
#undef LCB_CE
    const double oc0 = __dsub_rn(c0, o[0]), oc1 = __dsub_rn(c1, o[1]), oc2 = __dsub_rn(c2, o[2]);
    const double dd = __dadd_rn(__dadd_rn(__dmul_rn(d[0], d[0]), __dmul_rn(d[1], d[1])), __dmul_rn(d[2], d[2]));
    const double b = __dadd_rn(__dadd_rn(__dmul_rn(oc0, d[0]), __dmul_rn(oc1, d[1])), __dmul_rn(oc2, d[2]));
    const double ococ = __dadd_rn(__dadd_rn(__dmul_rn(oc0, oc0), __dmul_rn(oc1, oc1)), __dmul_rn(oc2, oc2));
    const double disc = __dsub_rn(__dmul_rn(b, b), __dmul_rn(dd, __dsub_rn(ococ, __dmul_rn(c3, c3))));
    if (!(disc >= 0.0)) return false;
    t_out = __double2float_rn(__ddiv_rn(__dsub_rn(b, __dsqrt_rn(disc)), dd)); u_out = __double2float_rn(ue);
    return true;
}
__device__ __forceinline__ bool refine_curve_hit(const RaySetup &r, const float4 *__restrict__ coef, int end, float &t_io, float &u_io) {
    float4 a[4] = {__ldg(coef), __ldg(coef + 1), __ldg(coef + 2), __ldg(coef + 3)};
    double t = (double)t_io, u = (double)u_io;
    const double o[3] = {(double)r.ox, (double)r.oy, (double)r.oz}, d[3] = {(double)r.dx, (double)r.dy, (double)r.dz};
    double step_t = INFINITY, step_u = INFINITY, qd = 0.0;
    bool settled = true;
    for (int it = 0; it < 6; it++) {
        double c[4], c1[4], c2[4];
#define LCB_EV(K, C)                                                                                                                           \
        {                                                                                                                                      \
            const double a0 = (double)a[0].C, a1 = (double)a[1].C, a2 = (double)a[2].C, a3 = (double)a[3].C;                                     \
            c[K] = __dadd_rn(__dmul_rn(__dadd_rn(__dmul_rn(__dadd_rn(__dmul_rn(a0, u), a1), u), a2), u), a3);                                   \
            c1[K] = __dadd_rn(__dmul_rn(__dadd_rn(__dmul_rn(__dmul_rn(3.0, a0), u), __dmul_rn(2.0, a1)), u), a2);                               \
            c2[K] = __dadd_rn(__dmul_rn(__dmul_rn(6.0, a0), u), __dmul_rn(2.0, a1));                                                             \
        }
        LCB_EV(0, x) LCB_EV(1, y) LCB_EV(2, z) LCB_EV(3, w)
#undef LCB_EV
        const double q0 = __dsub_rn(__dadd_rn(o[0], __dmul_rn(t, d[0])), c[0]), q1 = __dsub_rn(__dadd_rn(o[1], __dmul_rn(t, d[1])), c[1]),
                     q2 = __dsub_rn(__dadd_rn(o[2], __dmul_rn(t, d[2])), c[2]);
#define LCB_DOT3(X0, X1, X2, Y0, Y1, Y2) __dadd_rn(__dadd_rn(__dmul_rn(X0, Y0), __dmul_rn(X1, Y1)), __dmul_rn(X2, Y2))
        const double qq = LCB_DOT3(q0, q1, q2, q0, q1, q2);
        qd = LCB_DOT3(q0, q1, q2, d[0], d[1], d[2]);
        const double qc1 = LCB_DOT3(q0, q1, q2, c1[0], c1[1], c1[2]);
        const double qc2 = LCB_DOT3(q0, q1, q2, c2[0], c2[1], c2[2]);
        const double dc1 = LCB_DOT3(d[0], d[1], d[2], c1[0], c1[1], c1[2]);
        const double c1c1 = LCB_DOT3(c1[0], c1[1], c1[2], c1[0], c1[1], c1[2]);
#undef LCB_DOT3
        const double F1 = __dsub_rn(qq, __dmul_rn(c[3], c[3]));
        const double F2 = __dadd_rn(qc1, __dmul_rn(c[3], c1[3]));
        const double J11 = __dmul_rn(2.0, qd), J12 = __dmul_rn(-2.0, F2), J21 = dc1;
        const double J22 = __dadd_rn(__dadd_rn(__dsub_rn(qc2, c1c1), __dmul_rn(c1[3], c1[3])), __dmul_rn(c[3], c2[3]));
        const double det = __dsub_rn(__dmul_rn(J11, J22), __dmul_rn(J12, J21));
        if (det == 0.0) { settled = false; break; }
        step_t = __ddiv_rn(__dsub_rn(__dmul_rn(F1, J22), __dmul_rn(J12, F2)), det);
        step_u = __ddiv_rn(__dsub_rn(__dmul_rn(J11, F2), __dmul_rn(J21, F1)), det);
        t = __dsub_rn(t, step_t);
        u = __dsub_rn(u, step_u);
        if (!(fabs(u) <= 4.0)) { settled = false; break; }
        if (fabs(step_t) <= __dmul_rn(1e-13, __dadd_rn(fabs(t), 1.0)) && fabs(step_u) <= 1e-13) break;
    }
    if (settled && !(fabs(step_t) <= __dmul_rn(1e-7, __dadd_rn(fabs(t), 1.0)) && fabs(step_u) <= 1e-7)) settled = false;
    bool found = false;
    float tb = 0.f, ub = 0.f;
    if (settled && u >= 0.0 && u <= 1.0 && qd < 0.0) { found = true; tb = __double2float_rn(t); ub = __double2float_rn(u); }
    if (end != 0) {
        float te, ue;
        if (curve_end_sphere(o, d, a, end < 0 ? 0.0 : 1.0, te, ue) && (!found || te < tb)) { found = true; tb = te; ub = ue; }
    }
    if (!found) return false;
    t_io = tb; u_io = ub;
    return true;
}

__device__ __forceinline__ float dot3_rn(float ax, float ay, float az, float bx, float by, float bz) {
    return __fmaf_rn(ax, bx, __fmaf_rn(ay, by, __fmul_rn(az, bz)));
}
// Canonical ray / rounded cone: the smallest t in (tmin, tmax] among the roots of the lateral surface that lie between the two
// tangent circles and the entries into the two end spheres (after I. Quilez' rounded-cone intersector, extended to unnormalised
// directions and both lateral roots).  The quadratics are formed about the point of the ray nearest to sphere A, so that their
// cancellation error scales with the piece and not with the distance the ray has travelled.  s = axis parameter of the sphere of
// the sweep that the hit point lies on.
__device__ __forceinline__ bool canonical_cone(const RaySetup &r, float tmin, float tmax, const float4 A, const float4 B, float &t_out, float &s_out) {
    const float ra = A.w, rb = B.w;
    const float dd = dot3_rn(r.dx, r.dy, r.dz, r.dx, r.dy, r.dz);
    const float t0 = __fdiv_rn(dot3_rn(__fsub_rn(A.x, r.ox), __fsub_rn(A.y, r.oy), __fsub_rn(A.z, r.oz), r.dx, r.dy, r.dz), dd);
    const float ox = __fmaf_rn(t0, r.dx, r.ox), oy = __fmaf_rn(t0, r.dy, r.oy), oz = __fmaf_rn(t0, r.dz, r.oz);
    const float bax = __fsub_rn(B.x, A.x), bay = __fsub_rn(B.y, A.y), baz = __fsub_rn(B.z, A.z);
    const float oax = __fsub_rn(ox, A.x), oay = __fsub_rn(oy, A.y), oaz = __fsub_rn(oz, A.z);
    const float obx = __fsub_rn(ox, B.x), oby = __fsub_rn(oy, B.y), obz = __fsub_rn(oz, B.z);
    const float rr = __fsub_rn(ra, rb);
    const float m0 = dot3_rn(bax, bay, baz, bax, bay, baz), m1 = dot3_rn(bax, bay, baz, oax, oay, oaz), m2 = dot3_rn(bax, bay, baz, r.dx, r.dy, r.dz);
    const float m3 = dot3_rn(r.dx, r.dy, r.dz, oax, oay, oaz), m5 = dot3_rn(oax, oay, oaz, oax, oay, oaz);
    const float m6 = dot3_rn(obx, oby, obz, r.dx, r.dy, r.dz), m7 = dot3_rn(obx, oby, obz, obx, oby, obz);
    const float d2 = __fmaf_rn(-rr, rr, m0);
    bool found = false;
    float tb = 0.f, sb = 0.f;
    if (d2 > 0.0f) {
        const float k2 = __fmaf_rn(d2, dd, -__fmul_rn(m2, m2));
        const float k1 = __fmaf_rn(d2, m3, __fmaf_rn(-m1, m2, __fmul_rn(__fmul_rn(m2, rr), ra)));
        const float k0 = __fmaf_rn(d2, m5, __fmaf_rn(-m1, m1, __fmaf_rn(__fmul_rn(m1, rr), __fmul_rn(ra, 2.0f), -__fmul_rn(m0, __fmul_rn(ra, ra)))));
        const float h = __fmaf_rn(k1, k1, -__fmul_rn(k0, k2));
        if (h >= 0.0f && k2 != 0.0f) {
            const float sq = __fsqrt_rn(h);
#pragma unroll
            for (int root = 0; root < 2; root++) {
                const float tl = __fdiv_rn(root == 0 ? __fsub_rn(-sq, k1) : __fsub_rn(sq, k1), k2);
                const float y = __fmaf_rn(tl, m2, __fmaf_rn(-ra, rr, m1));
                const float t = __fadd_rn(tl, t0);
                if (t > tmin && t <= tmax && y > 0.0f && y < d2 && (!found || t < tb)) { found = true; tb = t; sb = __fdiv_rn(y, d2); }
            }
        }
    }
    const float h1 = __fmaf_rn(m3, m3, -__fmul_rn(dd, __fmaf_rn(-ra, ra, m5)));
    if (h1 >= 0.0f) {
        const float t = __fadd_rn(__fdiv_rn(__fsub_rn(-m3, __fsqrt_rn(h1)), dd), t0);
        if (t > tmin && t <= tmax && (!found || t < tb)) { found = true; tb = t; sb = 0.0f; }
    }
    const float h2 = __fmaf_rn(m6, m6, -__fmul_rn(dd, __fmaf_rn(-rb, rb, m7)));
    if (h2 >= 0.0f) {
        const float t = __fadd_rn(__fdiv_rn(__fsub_rn(-m6, __fsqrt_rn(h2)), dd), t0);
        if (t > tmin && t <= tmax && (!found || t < tb)) { found = true; tb = t; sb = 1.0f; }
    }
    t_out = tb; s_out = sb;
    return found;
}

// ---- one ray, traced by the calling thread ---------------------------------------------------------------------------
// lc_trace_closest / lc_trace_any for device code: the same node test, the same canonical triangle arithmetic and the
// same tie rule as the batch kernel, as a plain per-thread loop with a local-memory stack.  Returns the hit in the
// reference's SurfaceHit convention (miss: inst = prim = ~0, bary = 0, t = ray.tmax); ANY returns only hit / no hit.
struct DeviceHit { uint32_t inst, prim; float u, v, t; uint32_t kind; };  // kind: 0 miss, 1 triangle, 2 procedural (HitType, rtx.rs:510-516)

// QUERY adds the RayQuery rules (AccelImpl::ray_query, cpu/accel.rs:582-800; batch form: trace.cu kQueryAll / kQueryAny):
// triangles of NON-opaque instances are candidates handed to `hook(inst, prim, u, v, t)` with the canonical fp32
// barycentrics; it returns bit 0 = commit (RayQueryCommitTriangle), bit 1 = terminate (RayQueryTerminate).  Triangles of
// opaque instances commit directly.  `first` ends the traversal at the first committed hit (RayQueryAny).
// Leaves of procedural instances (user AABBs) are candidates for `hook.procedural(inst, prim, t_far, t)`, which may commit with
// its own t (RayQueryCommitProcedural); it is accepted when tmin <= t < t_far (cpu/accel.rs:711-713), ties on t going to the
// lowest (inst, prim) like everywhere else.  Without QUERY procedural instances are not entered at all.
struct NoCandidateHook {
    __device__ __forceinline__ int triangle(uint32_t, uint32_t, float, float, float) const { return 1; }
    __device__ __forceinline__ int procedural(uint32_t, uint32_t, float, float &) const { return 0; }
};

// CURVES: curve instances (flags bit 3) are entered and their leaves tested with canonical_cone; a curve hit goes through the same
// commit rules as a triangle (opaque: commits; otherwise hook.triangle with bary = (u, -1): the reference routes curve candidates of
// a RayQuery to on_surface_hit as well, cpu/accel.rs:650-684).  Without CURVES those instances are skipped.  Lowered kernels get it
// from the module's curve_basis_set (the frontend records it per trace call, rtx.rs:780-806), like the OptiX backend does.
#ifndef LCB_CURVES
#define LCB_CURVES 0
#endif
template <bool ANY, bool QUERY, class Hook, bool CURVES = (LCB_CURVES != 0)>
__device__ __forceinline__ DeviceHit trace_one_impl(const AccelView &acc, const float4 ra, const float4 rb, uint32_t mask, bool first, Hook &hook) {
    DeviceHit h{kNone, kNone, 0.f, 0.f, rb.w, 0u};
    if (!acc.tlas_nodes) return h;
    uint2 stack[kTraversalStack];
    int sp = 0;
    RaySetup r;
    setup_world(r, ra, rb);
    const float tmin = ra.w, ray_tmax = rb.w;
    float tbest = rb.w;
    uint32_t cur_inst = kNone, hit_slot = 0;
    bool hit_curve = false;
    bool cur_opaque = true, cur_procedural = false, cur_curve = false, stop = false;
    const WideNode *nodes = acc.tlas_nodes;
    const PackedTri *tris = nullptr;
    uint2 G = make_uint2(0u, 0x80000000u), Gt = make_uint2(0u, 0u);
    for (;;) {
        if (Gt.y == 0u && (G.y & 0xff000000u) != 0u) {
            const uint32_t bit = 31u - __clz(G.y);
            G.y &= ~(1u << bit);
            const uint32_t slot = (bit - 24u) ^ r.octinv;
            const uint32_t rel = __popc(G.y & 0xffu & ((1u << slot) - 1u));
            const WideNode *node = nodes + (G.x + rel);
            if (G.y & 0xff000000u) stack[sp++] = G;
            uint32_t child_base, prim_base, imask;
            const uint32_t hits = intersect_node(node, r, tmin, tbest, child_base, prim_base, imask);
            G = make_uint2(child_base, (hits & 0xff000000u) | imask);
            Gt = make_uint2(prim_base, hits & 0x00ffffffu);
        }
        while (Gt.y != 0u) {
            const uint32_t bit = __ffs(Gt.y) - 1;
            Gt.y &= Gt.y - 1;
            if (QUERY && cur_procedural) {
                const uint32_t prim = __float_as_uint(__ldg(reinterpret_cast<const float4 *>(tris + (Gt.x + bit))).w);
                float t = 0.f;
                const int verdict = hook.procedural(cur_inst, prim, tbest, t);
                stop = (verdict & 2) != 0;
                if ((verdict & 1) && t >= tmin && (t < tbest || (t == tbest && h.inst != kNone && (cur_inst < h.inst || (cur_inst == h.inst && prim < h.prim))))) {
                    tbest = t; h.inst = cur_inst; h.prim = prim; h.kind = 2u; hit_slot = Gt.x + bit;
                    if (first) stop = true;
                }
                if (stop) { Gt.y = 0u; G.y = 0u; sp = 0; break; }
            } else if (CURVES && cur_curve) {
                const float4 *tp = reinterpret_cast<const float4 *>(tris + (Gt.x + bit));
                const float4 c0 = __ldg(tp), c1 = __ldg(tp + 1), c2 = __ldg(tp + 2);
                float t, sl;
                bool cone_hit = canonical_cone(r, tmin, ray_tmax, make_float4(c0.x, c0.y, c0.z, c1.w), make_float4(c1.x, c1.y, c1.z, c2.x), t, sl);
                float u = __fmaf_rn(sl, c2.z, c2.y);
                if (cone_hit && c2.z != 1.0f) {  // a piece of a cubic segment: the cone located the hit, the swept surface decides (refine_curve_hit)
                    const int end = c2.y == 0.0f ? -1 : (__fadd_rn(c2.y, c2.z) == 1.0f ? 1 : 0);
                    cone_hit = refine_curve_hit(r, reinterpret_cast<const float4 *>(tris + __float_as_uint(c2.w)), end, t, u) && t > tmin && t <= ray_tmax;
                }
                if (cone_hit) {
                    const uint32_t prim = __float_as_uint(c0.w);
                    bool commit = true;
                    if (QUERY && !cur_opaque) {
                        const int verdict = hook.triangle(cur_inst, prim, u, -1.0f, t);
                        commit = (verdict & 1) != 0; stop = (verdict & 2) != 0;
                    }
                    if (commit) {
                        if (ANY) { h.inst = cur_inst; h.prim = prim; h.t = t; h.kind = 1u; return h; }
                        // ties: lowest (inst, prim), then lowest u (two pieces of one segment meeting in a shared sphere)
                        const bool better = t < tbest || h.inst == kNone ||
                                            (t == tbest && (cur_inst < h.inst || (cur_inst == h.inst && (prim < h.prim || (prim == h.prim && hit_curve && u < h.u)))));
                        if (better) { tbest = t; h.inst = cur_inst; h.prim = prim; h.kind = 1u; h.u = u; h.v = -1.0f; hit_curve = true; }
                        if (QUERY && first) stop = true;
                    }
                    if (QUERY && stop) { Gt.y = 0u; G.y = 0u; sp = 0; break; }
                }
            } else if (cur_inst != kNone) {
                const float4 *tp = reinterpret_cast<const float4 *>(tris + (Gt.x + bit));
                const float4 v0 = __ldg(tp), v1 = __ldg(tp + 1), v2 = __ldg(tp + 2);
                float t, V, W, det;
                if (canonical_triangle(r, tmin, ray_tmax, v0, v1, v2, t, V, W, det)) {
                    const uint32_t prim = __float_as_uint(v0.w);
                    bool commit = true;
                    if (QUERY && !cur_opaque) {
                        const float rdet = __frcp_rn(det);
                        const int verdict = hook.triangle(cur_inst, prim, __fmul_rn(V, rdet), __fmul_rn(W, rdet), t);
                        commit = (verdict & 1) != 0; stop = (verdict & 2) != 0;
                    }
                    if (commit) {
                        if (ANY) { h.inst = cur_inst; h.prim = prim; h.t = t; h.kind = 1u; return h; }
                        const bool better = t < tbest || h.inst == kNone || (t == tbest && (cur_inst < h.inst || (cur_inst == h.inst && prim < h.prim)));
                        if (better) { tbest = t; h.inst = cur_inst; h.prim = prim; h.kind = 1u; hit_slot = Gt.x + bit; hit_curve = false; }
                        if (QUERY && first) stop = true;
                    }
                    if (QUERY && stop) { Gt.y = 0u; G.y = 0u; sp = 0; break; }
                }
            } else {
                const uint32_t inst = __ldg(acc.tlas_prims + Gt.x + bit);
                const float4 *rec = reinterpret_cast<const float4 *>(acc.instances + inst);
                const uint4 meta = __ldg(reinterpret_cast<const uint4 *>(rec) + 4);
                if ((meta.x & mask) != 0u && (QUERY || (meta.z & 4u) == 0u) && (CURVES || (meta.z & 8u) == 0u)) {
                    if (Gt.y) stack[sp++] = Gt;
                    if (G.y & 0xff000000u) stack[sp++] = G;
                    const uint4 ptrs = __ldg(reinterpret_cast<const uint4 *>(rec) + 3);
                    nodes = reinterpret_cast<const WideNode *>(((unsigned long long)ptrs.y << 32) | ptrs.x);
                    tris = reinterpret_cast<const PackedTri *>(((unsigned long long)ptrs.w << 32) | ptrs.z);
                    stack[sp++] = make_uint2(enter_instance(r, meta.z, rec) ? 1u : 0u, 0u);  // sentinel: below it lies world space (x = 1: same ray setup)
                    cur_inst = inst;
                    if (QUERY) { cur_opaque = (meta.z & 2u) != 0u; cur_procedural = (meta.z & 4u) != 0u; }
                    if (CURVES) cur_curve = (meta.z & 8u) != 0u;
                    G = make_uint2(0u, 0x80000000u);
                    Gt = make_uint2(0u, 0u);
                    break;
                }
            }
        }
        if (Gt.y == 0u && (G.y & 0xff000000u) == 0u) {
            bool done = false;
            for (;;) {
                if (sp == 0) { done = true; break; }
                const uint2 e = stack[--sp];
                if (e.y & 0xff000000u) { G = e; break; }
                if (e.y != 0u) { Gt = e; break; }
                cur_inst = kNone; cur_procedural = false; cur_curve = false; nodes = acc.tlas_nodes; tris = nullptr;
                if (sp == 0) { done = true; break; }
                if (e.x == 0u) setup_world(r, ra, rb);
            }
            if (done) break;
        }
    }
    if (!ANY && h.inst != kNone) h.t = tbest;
    if (!ANY && h.inst != kNone && h.kind == 1u && !hit_curve) {
        const float4 *rec = reinterpret_cast<const float4 *>(acc.instances + h.inst);
        const float4 m0 = __ldg(rec), m1 = __ldg(rec + 1), m2 = __ldg(rec + 2);
        const uint4 ptrs = __ldg(reinterpret_cast<const uint4 *>(rec) + 3);
        const float4 *tp = reinterpret_cast<const float4 *>(reinterpret_cast<const PackedTri *>(((unsigned long long)ptrs.w << 32) | ptrs.z) + hit_slot);
        const float4 v0 = __ldg(tp), v1 = __ldg(tp + 1), v2 = __ldg(tp + 2);
        transform_ray(r, ra, rb, m0, m1, m2);
        if (!refine_bary(r, v0, v1, v2, h.u, h.v)) {
            finish_setup(r);
            float t, V, W, det;
            if (canonical_triangle(r, -INFINITY, INFINITY, v0, v1, v2, t, V, W, det)) {
                const float rdet = __frcp_rn(det);
                h.u = __fmul_rn(V, rdet); h.v = __fmul_rn(W, rdet);
            }
        }
    }
    return h;
}

template <bool ANY>
__device__ __forceinline__ DeviceHit trace_one(const AccelView &acc, const float4 ra, const float4 rb, uint32_t mask) {
    NoCandidateHook hook;
    return trace_one_impl<ANY, false>(acc, ra, rb, mask, false, hook);
}


// ---- wavefront traversal for kernels that trace from their own threads -----------------------------------------------------------
// The persistent-thread form of a lowered DSL kernel (ir_lower.cpp "wavefront lowering"): the kernel body is a resumable state machine,
// a trace call parks the lane's ray here (wave_begin) and yields, and the warp then runs the SAME if-if loop with postponing as the batch
// kernel k_trace (trace.cu) over the lanes that are parked — all of them converged in the wide node step, whatever point of the user
// code each came from.  The loop hands control back (wave_traverse returns) as soon as `yield_min` lanes could make progress in user
// code (their traversal ended, or they are waiting for a new work item), so finished lanes are refilled with new rays while the others
// keep their traversal state in registers — the work refill of k_trace, with the kernel body in the place of the ray fetch.
// Stack: kWaveSmemStack levels per thread in shared memory laid out [level][thread], deeper levels in the caller's local array.
// The world-space ray is parked in shared memory too (it is needed again when an instance is left and for the barycentrics).
constexpr int kWaveThreads = 128;     // CTA size of a wavefront-lowered kernel (4 independent warps)
constexpr int kWaveSmemStack = 16;
constexpr int kWaveLocalStack = kTraversalStack - kWaveSmemStack;
constexpr int kWaveChunk = 128;       // work items taken per global atomic
// Postponed primitive groups are optional pushes; the mandatory ones are one node group per level plus three per instance entry
// (<= 2 * kMaxWideDepth + 3).  Postponing stops where the two together could exceed the stack.
constexpr int kPostponeLimit = kTraversalStack - (2 * kMaxWideDepth + 3) - 1;
static_assert(kPostponeLimit >= 24, "traversal stack too small for postponing");

struct WaveShared {
    uint2 stack[kWaveSmemStack * kWaveThreads];
    float4 ray[2 * kWaveThreads];  // [0..T): origin | tmin, [T..2T): direction | tmax
};

struct WaveLane {
    RaySetup r;
    float tmin, tbest;  // tbest starts as the ray's tmax and only shrinks: it is also the upper end of the interval every triangle is tested against
    uint32_t hit_inst, hit_prim, hit_slot, cur_inst, mask;
    const WideNode *nodes;
    const PackedTri *tris;
    uint2 G, Gt;
    int sp;
    bool any;
};

// park one ray: returns false when there is nothing to traverse (empty accel) — the lane's result is then already a miss
__device__ __forceinline__ bool wave_begin(WaveLane &w, WaveShared &S, const AccelView &acc, const float4 ra, const float4 rb, uint32_t mask, bool any) {
    S.ray[threadIdx.x] = ra; S.ray[kWaveThreads + threadIdx.x] = rb;
    setup_world(w.r, ra, rb);
    w.tmin = ra.w; w.tbest = rb.w;
    w.hit_inst = kNone; w.hit_prim = kNone; w.hit_slot = 0u;
    w.cur_inst = kNone; w.nodes = acc.tlas_nodes; w.tris = nullptr;
    w.sp = 0; w.mask = mask; w.any = any;
    w.G = make_uint2(0u, acc.tlas_nodes ? 0x80000000u : 0u);
    w.Gt = make_uint2(0u, 0u);
    return acc.tlas_nodes != nullptr;
}

// `state` per lane: kWaveTraversing lanes take part; a lane whose ray is finished becomes kWaveReady.  `m_dead`: lanes that will never
// have anything to do again (warp-uniform).  Returns when no lane traverses any more or when yield_min lanes are ready.
constexpr int kWaveNeedsWork = 0, kWaveReady = 1, kWaveTraversing = 2, kWaveDead = 3;
__device__ __forceinline__ void wave_traverse(WaveLane &w, int &state, const AccelView &acc, WaveShared &S, uint2 *l_stack, uint32_t m_dead, int yield_min) {
    uint2 *const my_stack = S.stack + threadIdx.x;
#define LCW_PUSH(E)                                                          \
    {                                                                        \
        if (w.sp < kWaveSmemStack) my_stack[w.sp * kWaveThreads] = (E);      \
        else l_stack[w.sp - kWaveSmemStack] = (E);                           \
        w.sp++;                                                              \
    }
    for (;;) {
        const uint32_t m_trav = __ballot_sync(0xffffffffu, state == kWaveTraversing);
        if (m_trav == 0u) break;
        if (32 - __popc(m_trav | m_dead) >= yield_min) break;
        const bool has_ray = state == kWaveTraversing;
        // ---- node step ----------------------------------------------------------------------------------------------------------
        if (has_ray && w.Gt.y == 0u && (w.G.y & 0xff000000u) != 0u) {
            const uint32_t bit = 31u - __clz(w.G.y);
            w.G.y &= ~(1u << bit);
            const uint32_t slot = (bit - 24u) ^ w.r.octinv;
            const uint32_t rel = __popc(w.G.y & 0xffu & ((1u << slot) - 1u));
            const WideNode *node = w.nodes + (w.G.x + rel);
            if (w.G.y & 0xff000000u) LCW_PUSH(w.G)
            uint32_t child_base, prim_base, imask;
            const uint32_t hits = intersect_node(node, w.r, w.tmin, w.tbest, child_base, prim_base, imask);
            w.G = make_uint2(child_base, (hits & 0xff000000u) | imask);
            w.Gt = make_uint2(prim_base, hits & 0x00ffffffu);
        }
        // ---- primitive step: one triangle (or one instance entry) -------------------------------------------------------------------
        if (has_ray && w.Gt.y != 0u) {
            const uint32_t bit = __ffs(w.Gt.y) - 1;
            w.Gt.y &= w.Gt.y - 1;
            if (w.cur_inst != kNone) {
                const float4 *tp = reinterpret_cast<const float4 *>(w.tris + (w.Gt.x + bit));
                const U8 t01 = ldg256(tp);
                const float4 v2 = __ldg(tp + 2);
                const float4 v0 = make_float4(__uint_as_float(t01.v[0]), __uint_as_float(t01.v[1]), __uint_as_float(t01.v[2]), __uint_as_float(t01.v[3]));
                const float4 v1 = make_float4(__uint_as_float(t01.v[4]), __uint_as_float(t01.v[5]), __uint_as_float(t01.v[6]), 0.f);
                float t, V, W, det;
                // (tmin, tbest]: a hit beyond the best one so far would lose anyway; ties at tbest still reach the (inst, prim) rule below
                if (canonical_triangle(w.r, w.tmin, w.tbest, v0, v1, v2, t, V, W, det)) {
                    const uint32_t prim = __float_as_uint(v0.w);
                    if (w.any) {
                        w.hit_inst = w.cur_inst; w.Gt.y = 0u; w.G.y = 0u; w.sp = 0;  // retires in the tail below
                    } else {
                        const bool better = t < w.tbest || w.hit_inst == kNone ||
                                            (t == w.tbest && (w.cur_inst < w.hit_inst || (w.cur_inst == w.hit_inst && prim < w.hit_prim)));
                        if (better) { w.tbest = t; w.hit_inst = w.cur_inst; w.hit_prim = prim; w.hit_slot = w.Gt.x + bit; }
                    }
                }
            } else {
                const uint32_t inst = __ldg(acc.tlas_prims + w.Gt.x + bit);
                const float4 *rec = reinterpret_cast<const float4 *>(acc.instances + inst);
                const uint4 meta = __ldg(reinterpret_cast<const uint4 *>(rec) + 4);  // visibility, user_id, flags, pad
                // procedural (bit 2) and curve (bit 3) instances are not entered: trace_closest / trace_any without a RayQuery and
                // without a curve basis skip them (trace_one_impl<ANY, false, ..., false>)
                if ((meta.x & w.mask) != 0u && (meta.z & 12u) == 0u) {
                    if (w.Gt.y) LCW_PUSH(w.Gt)
                    if (w.G.y & 0xff000000u) LCW_PUSH(w.G)
                    const uint4 ptrs = __ldg(reinterpret_cast<const uint4 *>(rec) + 3);
                    w.nodes = reinterpret_cast<const WideNode *>(((unsigned long long)ptrs.y << 32) | ptrs.x);
                    w.tris = reinterpret_cast<const PackedTri *>(((unsigned long long)ptrs.w << 32) | ptrs.z);
                    LCW_PUSH(make_uint2(enter_instance(w.r, meta.z, rec) ? 1u : 0u, 0u))  // sentinel: below it lies world space (x = 1: same ray setup)
                    w.cur_inst = inst;
                    w.G = make_uint2(0u, 0x80000000u);
                    w.Gt = make_uint2(0u, 0u);
                }
            }
            if (w.Gt.y != 0u && (w.G.y & 0xff000000u) != 0u && w.sp < kPostponeLimit) { LCW_PUSH(w.Gt) w.Gt.y = 0u; }
        }
        // ---- tail: pop the next group, or finish ---------------------------------------------------------------------------------------
        if (has_ray && w.Gt.y == 0u && (w.G.y & 0xff000000u) == 0u) {
            for (;;) {
                if (w.sp == 0) { state = kWaveReady; break; }
                --w.sp;
                const uint2 e = w.sp < kWaveSmemStack ? my_stack[w.sp * kWaveThreads] : l_stack[w.sp - kWaveSmemStack];
                if (e.y & 0xff000000u) { w.G = e; break; }
                if (e.y != 0u) { w.Gt = e; break; }
                w.cur_inst = kNone; w.nodes = acc.tlas_nodes; w.tris = nullptr;  // sentinel: the instance is exhausted
                if (w.sp == 0) { state = kWaveReady; break; }
                if (e.x == 0u) setup_world(w.r, S.ray[threadIdx.x], S.ray[kWaveThreads + threadIdx.x]);
            }
        }
    }
#undef LCW_PUSH
}

// result of a finished closest-hit traversal in the SurfaceHit convention, barycentrics as trace_one_impl reports them (refine_bary)
__device__ __forceinline__ DeviceHit wave_closest_result(const WaveLane &w, const WaveShared &S, const AccelView &acc) {
    const float4 ra = S.ray[threadIdx.x], rb = S.ray[kWaveThreads + threadIdx.x];
    DeviceHit h{w.hit_inst, w.hit_prim, 0.f, 0.f, rb.w, 0u};
    if (w.hit_inst == kNone) return h;
    h.t = w.tbest; h.kind = 1u;
    const float4 *rec = reinterpret_cast<const float4 *>(acc.instances + w.hit_inst);
    const float4 m0 = __ldg(rec), m1 = __ldg(rec + 1), m2 = __ldg(rec + 2);
    const uint4 ptrs = __ldg(reinterpret_cast<const uint4 *>(rec) + 3);
    const float4 *tp = reinterpret_cast<const float4 *>(reinterpret_cast<const PackedTri *>(((unsigned long long)ptrs.w << 32) | ptrs.z) + w.hit_slot);
    const float4 v0 = __ldg(tp), v1 = __ldg(tp + 1), v2 = __ldg(tp + 2);
    RaySetup r;
    transform_ray(r, ra, rb, m0, m1, m2);
    if (!refine_bary(r, v0, v1, v2, h.u, h.v)) {
        finish_setup(r);
        float t, V, W, det;
        if (canonical_triangle(r, -INFINITY, INFINITY, v0, v1, v2, t, V, W, det)) {
            const float rdet = __frcp_rn(det);
            h.u = __fmul_rn(V, rdet); h.v = __fmul_rn(W, rdet);
        }
    }
    return h;
}

// Warp-local pool of work items (the ray pool of k_trace): lanes in state kWaveNeedsWork receive consecutive items, topped up with one
// global atomic per kWaveChunk items; when the dispatch is exhausted they become kWaveDead.
struct WavePool { unsigned long long next, end; bool exhausted; };
__device__ __forceinline__ void wave_fetch(WavePool &pool, int &state, unsigned long long &item, unsigned long long total, unsigned long long *counter) {
    const uint32_t lane = threadIdx.x & 31u;
    for (;;) {
        const uint32_t m_need = __ballot_sync(0xffffffffu, state == kWaveNeedsWork);
        if (m_need == 0u) return;
        if (pool.next == pool.end) {
            if (pool.exhausted) { if (state == kWaveNeedsWork) state = kWaveDead; return; }
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(counter, (unsigned long long)kWaveChunk);
            base = __shfl_sync(0xffffffffu, base, 0);
            if (base >= total) { pool.exhausted = true; continue; }
            pool.next = base; pool.end = base + kWaveChunk < total ? base + kWaveChunk : total;
        }
        const unsigned long long mine = pool.next + __popc(m_need & ((1u << lane) - 1u));
        if (state == kWaveNeedsWork && mine < pool.end) { item = mine; state = kWaveReady; }
        const unsigned long long adv = pool.next + __popc(m_need);
        pool.next = adv < pool.end ? adv : pool.end;
    }
}

}  // namespace lcb
