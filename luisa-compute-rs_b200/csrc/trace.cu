// trace.cu — persistent-thread two-level traversal of the 8-wide quantised BVH.
//
// Replaces rtcIntersect1 / rtcOccluded1 behind AccelImpl::trace_closest / trace_any
// (cpu/accel.rs:449-535) for whole buffers of rays.
//
// Execution model: one ray per lane; warps are persistent and refill idle lanes from a
// warp-local pool of ray indices that is topped up with ONE global atomic per 128 rays
// (warp-synchronous work distribution, ballots only).  Every trip of the main loop is
// "if-if" with postponing (after Aila & Laine 2009 and Ylitie et al. 2017): all lanes that own a
// node group pop one child and test its 8 quantised child boxes, then the lanes that hold
// primitives test ONE of them; a group that is not exhausted by that is postponed (pushed) if
// the lane still has node work, so that the warp returns to the wide node step together.
// A node is one 128-byte line fetched as 4 x LDG.256 (sm_100's 256-bit loads halve the L1TEX
// wavefronts of the divergent node fetch; near/far planes are then picked with register selects).  The traversal stack holds one
// (child_base, hit mask) group per level: the first kSmemStack levels live in shared memory
// laid out [level][thread] (bank = f(thread) only, so pushes and pops at divergent depths are
// conflict-free), deeper levels spill to local memory.  Leaving an instance re-reads the
// 32-byte ray instead of keeping the world-space setup on the stack.
//
// The per-triangle arithmetic is the canonical fp32 sequence documented in DESIGN.md §3 and
// restated independently in oracle/oracle.c: every operation is an explicit round-to-nearest
// intrinsic so that nvcc cannot contract or reorder it.  Box culling is conservative with
// respect to that arithmetic (planes padded by 2^-20 of the L-inf distance to the node), so
// results do not depend on the tree nor on the schedule.
#include "build.cuh"
#include <cstdio>
#include <cstdlib>

namespace lcb {

namespace {

constexpr uint32_t kFull = 0xffffffffu;
constexpr uint32_t kNone = 0xffffffffu;
constexpr int kTraceThreads = 128;
constexpr int kChunk = 128;  // ray indices fetched per global atomic
#ifndef LCB_TRACE_MIN_BLOCKS
#define LCB_TRACE_MIN_BLOCKS 7
#endif
constexpr int kSmemStack = 16;                               // stack levels held in shared memory ([level][thread])
constexpr int kLocalStack = kTraversalStack - kSmemStack;    // deeper levels spill to local memory (never on the bench scenes)

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

// Tests the 8 children of one node.  Returns the hit mask: bits 24..31 internal children in
// traversal priority order for this ray's octant, bits 0..23 leaf primitives.
__device__ __forceinline__ uint32_t intersect_node(const WideNode *__restrict__ node, const RaySetup &r, float tmin, float tmax,
                                                   uint32_t &child_base, uint32_t &prim_base, uint32_t &imask) {
    const uint4 *p = reinterpret_cast<const uint4 *>(node);
    const U8 H = ldg256(p), X = ldg256(p + 2), Y = ldg256(p + 4), Z = ldg256(p + 6);
    const uint4 n0 = make_uint4(H.v[0], H.v[1], H.v[2], H.v[3]), n1 = make_uint4(H.v[4], H.v[5], H.v[6], H.v[7]);
    const bool neg_x = (r.octinv & 1u) == 0, neg_y = (r.octinv & 2u) == 0, neg_z = (r.octinv & 4u) == 0;
    // near/far plane vectors by direction sign (words 0..3 = lower planes, 4..7 = upper planes of the 8 slots)
#define LCB_NEAR(V, NEG) make_uint4(NEG ? V.v[4] : V.v[0], NEG ? V.v[5] : V.v[1], NEG ? V.v[6] : V.v[2], NEG ? V.v[7] : V.v[3])
#define LCB_FAR(V, NEG) make_uint4(NEG ? V.v[0] : V.v[4], NEG ? V.v[1] : V.v[5], NEG ? V.v[2] : V.v[6], NEG ? V.v[3] : V.v[7])
    const uint4 qnx = LCB_NEAR(X, neg_x), qfx = LCB_FAR(X, neg_x);
    const uint4 qny = LCB_NEAR(Y, neg_y), qfy = LCB_FAR(Y, neg_y);
    const uint4 qnz = LCB_NEAR(Z, neg_z), qfz = LCB_FAR(Z, neg_z);
#undef LCB_NEAR
#undef LCB_FAR
    child_base = n1.x; prim_base = n1.y; imask = n0.w >> 24;
    const float sclx = __uint_as_float((n0.w & 0xffu) << 23), scly = __uint_as_float((n0.w & 0xff00u) << 15), sclz = __uint_as_float((n0.w & 0xff0000u) << 7);
    const float rx = __uint_as_float(n0.x) - r.ox, ry = __uint_as_float(n0.y) - r.oy, rz = __uint_as_float(n0.z) - r.oz;
    // conservative padding: 2^-20 of the L-inf distance from the ray origin to the far side of the node frame
    const float R = fmaxf(fmaxf(fabsf(rx) + 65535.0f * sclx, fabsf(ry) + 65535.0f * scly), fabsf(rz) + 65535.0f * sclz);
    const float pad = R * (1.0f / 1048576.0f);
    const float ax = sclx * r.ix, ay = scly * r.iy, az = sclz * r.iz;
    const float cx = rx * r.ix, cy = ry * r.iy, cz = rz * r.iz;
    const float px = fmaf(pad, fabsf(r.ix), fabsf(ax)), py = fmaf(pad, fabsf(r.iy), fabsf(ay)), pz = fmaf(pad, fabsf(r.iz), fabsf(az));
    const float bnx = fmaf(-8388608.0f, ax, cx - px), bfx = fmaf(-8388608.0f, ax, cx + px);
    const float bny = fmaf(-8388608.0f, ay, cy - py), bfy = fmaf(-8388608.0f, ay, cy + py);
    const float bnz = fmaf(-8388608.0f, az, cz - pz), bfz = fmaf(-8388608.0f, az, cz + pz);
    // hit-mask construction on 4 meta bytes at a time (after Ylitie et al. 2017): per child only a byte extract,
    // a shift and a select remain.  Internal children (low 5 bits >= 24) get their bit index XORed with the
    // ray octant so that __clz order is front-to-back order.
    const uint32_t oct4 = r.octinv * 0x01010101u;
    const uint32_t inner_lo = sign_extend_bytes((n1.z & (n1.z << 1) & 0x10101010u) << 3);  // 0xff per internal child
    const uint32_t inner_hi = sign_extend_bytes((n1.w & (n1.w << 1) & 0x10101010u) << 3);
    const uint32_t idx_lo = (n1.z ^ (oct4 & inner_lo)) & 0x1f1f1f1fu, idx_hi = (n1.w ^ (oct4 & inner_hi)) & 0x1f1f1f1fu;
    const uint32_t bits_lo = (n1.z >> 5) & 0x07070707u, bits_hi = (n1.w >> 5) & 0x07070707u;
    uint32_t hits = 0;
#define LCB_CHILD(WORD, CONV, BITS, IDX, SHIFT)                                                              \
    {                                                                                                        \
        const float tn = fmaxf(fmaxf(fmaf(CONV(qnx.WORD), ax, bnx), fmaf(CONV(qny.WORD), ay, bny)),          \
                               fmaxf(fmaf(CONV(qnz.WORD), az, bnz), tmin));                                  \
        const float tf = fminf(fminf(fmaf(CONV(qfx.WORD), ax, bfx), fmaf(CONV(qfy.WORD), ay, bfy)),          \
                               fminf(fmaf(CONV(qfz.WORD), az, bfz), tmax));                                  \
        const uint32_t b = ((BITS >> SHIFT) & 0xffu) << ((IDX >> SHIFT) & 31u);                              \
        hits |= tn <= tf ? b : 0u;                                                                           \
    }
    LCB_CHILD(x, q16_lo, bits_lo, idx_lo, 0)
    LCB_CHILD(x, q16_hi, bits_lo, idx_lo, 8)
    LCB_CHILD(y, q16_lo, bits_lo, idx_lo, 16)
    LCB_CHILD(y, q16_hi, bits_lo, idx_lo, 24)
    LCB_CHILD(z, q16_lo, bits_hi, idx_hi, 0)
    LCB_CHILD(z, q16_hi, bits_hi, idx_hi, 8)
    LCB_CHILD(w, q16_lo, bits_hi, idx_hi, 16)
    LCB_CHILD(w, q16_hi, bits_hi, idx_hi, 24)
#undef LCB_CHILD
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

// Scheduling knob of the traversal loop (LC_B200_TRACE_TUNE = "fetch_min").
struct TraceTune {
    int fetch_min;      // refill when at least this many lanes are idle
};

// What a launch computes.  kQueryAll / kQueryAny are the batch forms of RayQuery (AccelImpl::ray_query,
// cpu/accel.rs:582-800; frontend rtx.rs:672-756): triangles of OPAQUE instances commit like in kClosest, candidates of
// NON-opaque instances go through the candidate hook (the on_surface_hit callback) and only count when it commits.
enum TraceMode : int { kClosest = 0, kAny = 1, kQueryAll = 2, kQueryAny = 3 };

// The candidate hook of the batch RayQuery: a pure function of the candidate (so the committed hit does not depend on
// traversal order).  A lowered DSL kernel would pass its own callable here.  Mirrored by oracle.c `filter_accept`.
__device__ __forceinline__ bool candidate_commits(const CandidateFilter &f, uint32_t inst, uint32_t prim, float u, float v) {
    switch (f.kind) {
        case 0: return true;                                       // commit every candidate
        case 1: {                                                  // examples/ray_query.rs:148-162: |uvw.xy|, |uvw.yz|, |uvw.xz| < r
            const float w = __fsub_rn(__fsub_rn(1.0f, u), v);      // uvw = (1-u-v, u, v)
            const float r2 = __fmul_rn(f.radius, f.radius);
            const float xy = __fmaf_rn(w, w, __fmul_rn(u, u)), yz = __fmaf_rn(u, u, __fmul_rn(v, v)), xz = __fmaf_rn(w, w, __fmul_rn(v, v));
            return xy < r2 && yz < r2 && xz < r2;
        }
        case 2: {                                                  // per-primitive cut-out bits: bit (first_bit[inst] + prim)
            const uint32_t b = __ldg(f.first_bit + inst) + prim;
            return (__ldg(f.bits + (b >> 5)) >> (b & 31u)) & 1u;
        }
        default: return false;                                     // reject every candidate
    }
}

template <int MODE, bool COUNTERS>
__global__ void __launch_bounds__(kTraceThreads, LCB_TRACE_MIN_BLOCKS) k_trace(AccelView acc, const float4 *__restrict__ rays, void *__restrict__ out,
                                                                               unsigned long long count, uint32_t mask, unsigned long long *work_counter,
                                                                               TraceCounters *ctr, TraceTune tune, const uint32_t *__restrict__ order,
                                                                               CandidateFilter filter) {
    constexpr bool ANY = MODE == kAny;
    constexpr bool QUERY = MODE == kQueryAll || MODE == kQueryAny;
    constexpr bool FIRST = MODE == kAny || MODE == kQueryAny;  // stop at the first committed hit
    bool cur_opaque = true;
    __shared__ uint2 s_stack[kSmemStack * kTraceThreads];
    uint2 l_stack[kLocalStack];
    const uint32_t lane = threadIdx.x & 31, lt_mask = (1u << lane) - 1;
    uint2 *const my_stack = s_stack + threadIdx.x;

    // warp-uniform pool of ray indices
    unsigned long long pool_next = 0, pool_end = 0;
    bool exhausted = false;

    bool has_ray = false;
    unsigned long long ray_idx = 0;
    RaySetup r;
    float tmin = 0.f, tbest = 0.f, ray_tmax = 0.f;
    uint32_t hit_inst = kNone, hit_prim = kNone, hit_slot = 0;  // hit_slot: index of the winning PackedTri
    uint32_t cur_inst = kNone;
    const WideNode *nodes = acc.tlas_nodes;
    const PackedTri *tris = nullptr;
    uint2 G = make_uint2(0, 0), Gt = make_uint2(0, 0);
    int sp = 0;
    unsigned long long n_nodes = 0, n_tris = 0, n_inst = 0, n_rays = 0;

#define LCB_PUSH(E)                                                          \
    {                                                                        \
        if (sp < kSmemStack) my_stack[sp * kTraceThreads] = (E);             \
        else l_stack[sp - kSmemStack] = (E);                                 \
        sp++;                                                                \
    }

    for (;;) {
        // Invariant: every lane that owns a ray has pending work (a non-empty node group G or primitive group Gt).
        // ---- fetch: refill idle lanes from the warp-local pool ------------------------------------------------
        const uint32_t m_idle = __ballot_sync(kFull, !has_ray);
        if (m_idle != 0u) {
            const bool can_fetch = !(exhausted && pool_next == pool_end);
            if (!can_fetch) {
                if (m_idle == kFull) break;
            } else if (__popc(m_idle) >= tune.fetch_min || m_idle == kFull) {
                if (pool_next == pool_end) {
                    unsigned long long base = 0;
                    if (lane == 0) base = atomicAdd(work_counter, (unsigned long long)kChunk);
                    base = __shfl_sync(kFull, base, 0);
                    if (base >= count) exhausted = true;
                    else { pool_next = base; pool_end = base + kChunk < count ? base + kChunk : count; }
                }
                if (pool_next != pool_end) {
                    const unsigned long long mine = pool_next + __popc(m_idle & lt_mask);
                    if (!has_ray && mine < pool_end) {
                        ray_idx = order ? (unsigned long long)__ldg(order + mine) : mine;
                        const float4 ra = __ldg(rays + 2 * ray_idx), rb = __ldg(rays + 2 * ray_idx + 1);
                        setup_world(r, ra, rb);
                        tmin = ra.w; tbest = rb.w; ray_tmax = rb.w;
                        hit_inst = kNone; hit_prim = kNone; hit_slot = 0;
                        cur_inst = kNone; nodes = acc.tlas_nodes; tris = nullptr;
                        sp = 0;
                        G = make_uint2(0u, acc.tlas_nodes ? 0x80000000u : 0u);
                        Gt = make_uint2(0u, 0u);
                        has_ray = true;
                        if (COUNTERS) n_rays++;
                    }
                    const unsigned long long adv = pool_next + __popc(m_idle);
                    pool_next = adv < pool_end ? adv : pool_end;
                } else if (m_idle == kFull) break;  // the batch is exhausted and no lane owns a ray
            }
        }

        // ---- node step: pop the nearest child of the node group, test its 8 children ----------------------------
        if (has_ray && Gt.y == 0u && (G.y & 0xff000000u) != 0u) {
            const uint32_t bit = 31u - __clz(G.y);
            G.y &= ~(1u << bit);
            const uint32_t slot = (bit - 24u) ^ r.octinv;
            const uint32_t rel = __popc(G.y & 0xffu & ((1u << slot) - 1u));
            const WideNode *node = nodes + (G.x + rel);
            if (G.y & 0xff000000u) LCB_PUSH(G)
            uint32_t child_base, prim_base, imask;
            const uint32_t hits = intersect_node(node, r, tmin, tbest, child_base, prim_base, imask);
            if (COUNTERS) n_nodes++;
            G = make_uint2(child_base, (hits & 0xff000000u) | imask);
            Gt = make_uint2(prim_base, hits & 0x00ffffffu);
        }

        // ---- primitive step: one triangle (or one instance entry) per lane that holds a primitive group -----------
        if (has_ray && Gt.y != 0u) {
            {
                const uint32_t bit = __ffs(Gt.y) - 1;
                Gt.y &= Gt.y - 1;
                if (cur_inst != kNone) {
                    const float4 *tp = reinterpret_cast<const float4 *>(tris + (Gt.x + bit));
                    const U8 t01 = ldg256(tp);
                    const float4 v2 = __ldg(tp + 2);
                    const float4 v0 = make_float4(__uint_as_float(t01.v[0]), __uint_as_float(t01.v[1]), __uint_as_float(t01.v[2]), __uint_as_float(t01.v[3]));
                    const float4 v1 = make_float4(__uint_as_float(t01.v[4]), __uint_as_float(t01.v[5]), __uint_as_float(t01.v[6]), 0.f);
                    if (COUNTERS) n_tris++;
                    float t, V, W, det;
                    if (canonical_triangle(r, tmin, ray_tmax, v0, v1, v2, t, V, W, det)) {
                        const uint32_t prim = __float_as_uint(v0.w);
                        bool commit = true;
                        if (QUERY && !cur_opaque) {  // candidate hook sees the canonical fp32 barycentrics
                            const float rdet = __frcp_rn(det);
                            commit = candidate_commits(filter, cur_inst, prim, __fmul_rn(V, rdet), __fmul_rn(W, rdet));
                        }
                        if (commit) {
                            if (ANY) {
                                hit_inst = cur_inst; Gt.y = 0u; G.y = 0u; sp = 0;  // retires in the tail below
                            } else {
                                const bool better = t < tbest || hit_inst == kNone ||
                                                    (t == tbest && (cur_inst < hit_inst || (cur_inst == hit_inst && prim < hit_prim)));
                                if (better) { tbest = t; hit_inst = cur_inst; hit_prim = prim; hit_slot = Gt.x + bit; }
                                if (FIRST) { Gt.y = 0u; G.y = 0u; sp = 0; }
                            }
                        }
                    }
                } else {
                    // TLAS leaf: transform the ray and descend into the instance's BLAS
                    const uint32_t inst = __ldg(acc.tlas_prims + Gt.x + bit);
                    const float4 *rec = reinterpret_cast<const float4 *>(acc.instances + inst);
                    const uint4 meta = __ldg(reinterpret_cast<const uint4 *>(rec) + 4);  // visibility, user_id, flags, pad
                    if ((meta.x & mask) != 0u) {
                        if (Gt.y) LCB_PUSH(Gt)
                        if (G.y & 0xff000000u) LCB_PUSH(G)
                        LCB_PUSH(make_uint2(0u, 0u))  // sentinel: below it lies world space
                        const float4 m0 = __ldg(rec), m1 = __ldg(rec + 1), m2 = __ldg(rec + 2);
                        const uint4 ptrs = __ldg(reinterpret_cast<const uint4 *>(rec) + 3);
                        nodes = reinterpret_cast<const WideNode *>(((unsigned long long)ptrs.y << 32) | ptrs.x);
                        tris = reinterpret_cast<const PackedTri *>(((unsigned long long)ptrs.w << 32) | ptrs.z);
                        const float4 wo = make_float4(r.ox, r.oy, r.oz, 0.f), wd = make_float4(r.dx, r.dy, r.dz, 0.f);
                        setup_object(r, wo, wd, m0, m1, m2);
                        cur_inst = inst;
                        if (QUERY) cur_opaque = (meta.z & 2u) != 0u;
                        G = make_uint2(0u, 0x80000000u);
                        Gt = make_uint2(0u, 0u);
                        if (COUNTERS) n_inst++;
                    }
                }
            }
            // A lane whose group is not exhausted by this one test postpones the rest (pushes it) when it still has
            // node work: on incoherent rays, returning to the wide node step with the whole warp beats draining the
            // group with a few stragglers (tune sweep in profiles/r01_trace_tune_sweep.txt).
            if (Gt.y != 0u && (G.y & 0xff000000u) != 0u) { LCB_PUSH(Gt) Gt.y = 0u; }
        }

        // ---- tail: lanes that ran out of work pop the next group, or retire ------------------------------------------
        if (has_ray && Gt.y == 0u && (G.y & 0xff000000u) == 0u) {
            bool retire = false;
            for (;;) {
                if (sp == 0) { retire = true; break; }
                --sp;
                const uint2 e = sp < kSmemStack ? my_stack[sp * kTraceThreads] : l_stack[sp - kSmemStack];
                if (e.y & 0xff000000u) { G = e; break; }
                if (e.y != 0u) { Gt = e; break; }
                // sentinel: the instance is exhausted, back to world space
                cur_inst = kNone; nodes = acc.tlas_nodes; tris = nullptr;
                if (sp == 0) { retire = true; break; }  // nothing left in the TLAS either: skip the re-setup
                const float4 ra = __ldg(rays + 2 * ray_idx), rb = __ldg(rays + 2 * ray_idx + 1);
                setup_world(r, ra, rb);
            }
            if (retire) {
                if (ANY) {
                    reinterpret_cast<uint32_t *>(out)[ray_idx] = hit_inst != kNone ? 1u : 0u;
                } else {
                    uint2 *o = reinterpret_cast<uint2 *>(out) + 3 * ray_idx;
                    o[0] = make_uint2(hit_inst, hit_prim);
                    // barycentrics are formed by k_refine; the pad word carries the winning PackedTri slot to it
                    o[2] = make_uint2(__float_as_uint(hit_inst != kNone ? tbest : ray_tmax), hit_slot);
                }
                has_ray = false;
            }
        }
    }
#undef LCB_PUSH
    if (COUNTERS) {
        atomicAdd(&ctr->nodes_visited, n_nodes);
        atomicAdd(&ctr->tris_tested, n_tris);
        atomicAdd(&ctr->instance_entries, n_inst);
        atomicAdd(&ctr->rays, n_rays);
    }
}

// Second pass of closest-hit queries: barycentrics of every hit, one thread per ray, fully converged.  The f64
// evaluation is the reported value (oracle.c refine_bary); if its determinant vanishes the canonical fp32 ones stand.
// COMMITTED selects the record layout: SurfaceHit {inst, prim, u, v, t, pad} or CommittedHit {inst, prim, u, v, hit_type, t}.
template <bool COMMITTED>
__global__ void __launch_bounds__(256) k_refine(AccelView acc, const float4 *__restrict__ rays, uint2 *__restrict__ hits, unsigned long long count) {
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const uint2 h0 = hits[3 * i], h2 = hits[3 * i + 2];
    if (h0.x == kNone) {
        hits[3 * i + 1] = make_uint2(0u, 0u);
        // a RayQuery that commits nothing leaves its zero-initialised CommittedHit (cpu_resource.h:320-331): hit_type Miss, t = 0
        hits[3 * i + 2] = COMMITTED ? make_uint2(0u /* HitType::Miss */, 0u) : make_uint2(h2.x, 0u);
        return;
    }
    const float4 *rec = reinterpret_cast<const float4 *>(acc.instances + h0.x);
    const float4 m0 = __ldg(rec), m1 = __ldg(rec + 1), m2 = __ldg(rec + 2);
    const uint4 ptrs = __ldg(reinterpret_cast<const uint4 *>(rec) + 3);
    const float4 *tp = reinterpret_cast<const float4 *>(reinterpret_cast<const PackedTri *>(((unsigned long long)ptrs.w << 32) | ptrs.z) + h2.y);
    const float4 ra = __ldg(rays + 2 * i), rb = __ldg(rays + 2 * i + 1);
    const float4 v0 = __ldg(tp), v1 = __ldg(tp + 1), v2 = __ldg(tp + 2);
    RaySetup r;
    transform_ray(r, ra, rb, m0, m1, m2);
    float u = 0.f, v = 0.f;
    if (!refine_bary(r, v0, v1, v2, u, v)) {
        finish_setup(r);
        float t, V, W, det;
        if (canonical_triangle(r, -INFINITY, INFINITY, v0, v1, v2, t, V, W, det)) {
            const float rdet = __frcp_rn(det);
            u = __fmul_rn(V, rdet); v = __fmul_rn(W, rdet);
        }
    }
    hits[3 * i + 1] = make_uint2(__float_as_uint(u), __float_as_uint(v));
    hits[3 * i + 2] = COMMITTED ? make_uint2(1u /* HitType::Triangle */, h2.x) : make_uint2(h2.x, 0u);
}

// ---- ray reordering ---------------------------------------------------------------------------------------------------
// Incoherent batches are traced in the order of a (origin cell, direction bin) key so that the lanes of a warp walk
// the same part of the tree: k_ray_keys builds the keys, the builder's onesweep sort orders them, k_trace fetches
// ray indices through the sorted permutation.  Results are written by original index and do not depend on the
// order (the arithmetic per ray is fixed), so this is purely a scheduling decision.
__device__ __forceinline__ uint32_t spread3(uint32_t x) {  // 10 bits -> every third bit
    x &= 0x3ffu;
    x = (x | (x << 16)) & 0x030000ffu;
    x = (x | (x << 8)) & 0x0300f00fu;
    x = (x | (x << 4)) & 0x030c30c3u;
    x = (x | (x << 2)) & 0x09249249u;
    return x;
}

__global__ void __launch_bounds__(256) k_ray_keys(const float4 *__restrict__ rays, uint32_t count, float3 lo, float3 inv_ext, int origin_bits, int dir_bits,
                                                  uint64_t *__restrict__ keys, uint32_t *__restrict__ vals) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const float4 a = __ldg(rays + 2 * (size_t)i), b = __ldg(rays + 2 * (size_t)i + 1);
    const float cells = (float)(1u << origin_bits), bins = (float)(1u << dir_bits);
    const uint32_t cmax = (1u << origin_bits) - 1u, bmax = (1u << dir_bits) - 1u;
    // NaN-safe clamps (fminf/fmaxf return the non-NaN operand)
    const uint32_t ox = min((uint32_t)fminf(fmaxf((a.x - lo.x) * inv_ext.x * cells, 0.f), 1023.f), cmax);
    const uint32_t oy = min((uint32_t)fminf(fmaxf((a.y - lo.y) * inv_ext.y * cells, 0.f), 1023.f), cmax);
    const uint32_t oz = min((uint32_t)fminf(fmaxf((a.z - lo.z) * inv_ext.z * cells, 0.f), 1023.f), cmax);
    const float m = fmaxf(fmaxf(fabsf(b.x), fabsf(b.y)), fabsf(b.z));
    const float s = m > 0.f ? 0.5f / m : 0.f;
    const uint32_t dx = min((uint32_t)fminf(fmaxf((b.x * s + 0.5f) * bins, 0.f), 1023.f), bmax);
    const uint32_t dy = min((uint32_t)fminf(fmaxf((b.y * s + 0.5f) * bins, 0.f), 1023.f), bmax);
    const uint32_t dz = min((uint32_t)fminf(fmaxf((b.z * s + 0.5f) * bins, 0.f), 1023.f), bmax);
    const uint32_t okey = spread3(ox) | (spread3(oy) << 1) | (spread3(oz) << 2);
    const uint32_t dkey = spread3(dx) | (spread3(dy) << 1) | (spread3(dz) << 2);
    // direction octant (the top Morton digit of the direction) leads, then the origin cell, then the finer direction bits
    const uint32_t dlow_bits = 3 * (dir_bits - 1);
    const uint32_t octant = dkey >> dlow_bits, dlow = dkey & ((1u << dlow_bits) - 1u);
    keys[i] = ((uint64_t)octant << (3 * origin_bits + dlow_bits)) | ((uint64_t)okey << dlow_bits) | dlow;
    vals[i] = i;
}

struct RaySortConfig { int enabled; unsigned long long min_count; int origin_bits, dir_bits; };
RaySortConfig ray_sort_config() {
    static RaySortConfig c = [] {
        RaySortConfig d{0, 1ull << 18, 5, 3};  // off by default: no gain on uniformly incoherent batches (profiles/r01_ray_sort_sweep.txt)
        if (const char *e = getenv("LC_B200_RAY_SORT")) sscanf(e, "%d,%llu,%d,%d", &d.enabled, &d.min_count, &d.origin_bits, &d.dir_bits);
        if (d.origin_bits < 1) d.origin_bits = 1;
        if (d.origin_bits > 10) d.origin_bits = 10;
        if (d.dir_bits < 1) d.dir_bits = 1;
        if (d.dir_bits > 6) d.dir_bits = 6;
        return d;
    }();
    return c;
}

TraceTune trace_tune() {
    static TraceTune t = [] {
        TraceTune d{6};
        if (const char *e = getenv("LC_B200_TRACE_TUNE")) sscanf(e, "%d", &d.fetch_min);
        return d;
    }();
    return t;
}

static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

template <int MODE, bool COUNTERS>
void launch(cudaStream_t s, const AccelView &a, const void *rays, void *out, uint64_t count, uint32_t mask, unsigned long long *work_counter,
            TraceCounters *ctr, LaunchCounter &lc, const CandidateFilter &filter = CandidateFilter{0, 0.f, nullptr, nullptr}) {
    constexpr bool ANY = MODE == kAny;
    static int blocks_per_sm = 0, sms = 0;
    if (!blocks_per_sm) {
        int dev = 0; cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, k_trace<MODE, COUNTERS>, kTraceThreads, 0);
        if (blocks_per_sm < 1) blocks_per_sm = 1;
    }
    unsigned long long want = (count + kTraceThreads - 1) / kTraceThreads;
    unsigned long long grid = (unsigned long long)sms * blocks_per_sm;
    if (grid > want) grid = want;
    if (grid == 0) return;
    // ray reordering for large batches
    const RaySortConfig rs = ray_sort_config();
    const uint32_t *order = nullptr;
    void *scratch = nullptr;
    if (rs.enabled && a.tlas_nodes && count >= rs.min_count && count < (1ull << 31)) {
        const uint32_t n = (uint32_t)count;
        const int key_bits = 3 * rs.origin_bits + 3 * rs.dir_bits, passes = (key_bits + 7) / 8;
        const size_t kb = align256((size_t)n * 8), vb = align256((size_t)n * 4), sb = align256(sort_scratch_bytes(n, passes));
        if (cudaMallocAsync(&scratch, 2 * kb + 2 * vb + sb, s) == cudaSuccess) {
            uint8_t *p = (uint8_t *)scratch;
            uint64_t *keys = (uint64_t *)p, *keys_alt = (uint64_t *)(p + kb);
            uint32_t *vals = (uint32_t *)(p + 2 * kb), *vals_alt = (uint32_t *)(p + 2 * kb + vb);
            float3 lo = make_float3(a.world_lo[0], a.world_lo[1], a.world_lo[2]), inv;
            inv.x = a.world_hi[0] > a.world_lo[0] ? 1.f / (a.world_hi[0] - a.world_lo[0]) : 0.f;
            inv.y = a.world_hi[1] > a.world_lo[1] ? 1.f / (a.world_hi[1] - a.world_lo[1]) : 0.f;
            inv.z = a.world_hi[2] > a.world_lo[2] ? 1.f / (a.world_hi[2] - a.world_lo[2]) : 0.f;
            k_ray_keys<<<(n + 255) / 256, 256, 0, s>>>(reinterpret_cast<const float4 *>(rays), n, lo, inv, rs.origin_bits, rs.dir_bits, keys, vals);
            lc.count++;
            const bool in_alt = sort_pairs(s, n, keys, vals, keys_alt, vals_alt, p + 2 * kb + 2 * vb, 0, passes, lc);
            order = in_alt ? vals_alt : vals;
        } else {
            (void)cudaGetLastError();  // no memory for the permutation: trace in submission order
            scratch = nullptr;
        }
    }
    cudaMemsetAsync(work_counter, 0, sizeof(unsigned long long), s);
    k_trace<MODE, COUNTERS><<<(unsigned)grid, kTraceThreads, 0, s>>>(a, reinterpret_cast<const float4 *>(rays), out, count, mask, work_counter, ctr, trace_tune(), order, filter);
    lc.count++;
    if (scratch) cudaFreeAsync(scratch, s);
    if (!ANY) {
        k_refine<MODE == kQueryAll || MODE == kQueryAny><<<(unsigned)((count + 255) / 256), 256, 0, s>>>(a, reinterpret_cast<const float4 *>(rays), reinterpret_cast<uint2 *>(out), count);
        lc.count++;
    }
}

}  // namespace

void trace_closest(cudaStream_t s, const AccelView &a, const void *rays, void *hits, uint64_t count, uint32_t mask, unsigned long long *work_counter,
                   TraceCounters *counters, LaunchCounter &lc) {
    if (counters) launch<kClosest, true>(s, a, rays, hits, count, mask, work_counter, counters, lc);
    else launch<kClosest, false>(s, a, rays, hits, count, mask, work_counter, nullptr, lc);
}

void trace_any(cudaStream_t s, const AccelView &a, const void *rays, uint32_t *occluded, uint64_t count, uint32_t mask, unsigned long long *work_counter,
               LaunchCounter &lc) {
    launch<kAny, false>(s, a, rays, occluded, count, mask, work_counter, nullptr, lc);
}

void ray_query(cudaStream_t s, const AccelView &a, const void *rays, void *committed_hits, uint64_t count, uint32_t mask, bool terminate_on_first,
               const CandidateFilter &filter, unsigned long long *work_counter, LaunchCounter &lc) {
    if (terminate_on_first) launch<kQueryAny, false>(s, a, rays, committed_hits, count, mask, work_counter, nullptr, lc, filter);
    else launch<kQueryAll, false>(s, a, rays, committed_hits, count, mask, work_counter, nullptr, lc, filter);
}

}  // namespace lcb
