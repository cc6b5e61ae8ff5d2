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
#include "trace_device.cuh"
#include <cstdio>
#include <cstdlib>

namespace lcb {

namespace {

constexpr uint32_t kFull = 0xffffffffu;
#ifndef LCB_TRACE_THREADS
#define LCB_TRACE_THREADS 128
#endif
constexpr int kTraceThreads = LCB_TRACE_THREADS;  // warps are independent (warp-local ray pools, no block barrier): any multiple of 32
constexpr int kChunk = 128;  // ray indices fetched per global atomic
#ifndef LCB_TRACE_MIN_BLOCKS
#define LCB_TRACE_MIN_BLOCKS 7  // 72 registers, no spills with the 96-byte node: 1410 Mrays/s on C3 against 1342 at 6 CTAs and 1202 at 5 (profiles/r02h_occupancy.txt)
#endif
#ifndef LCB_SMEM_STACK
#define LCB_SMEM_STACK 16
#endif
constexpr int kSmemStack = LCB_SMEM_STACK;                               // stack levels held in shared memory ([level][thread])
constexpr int kLocalStack = kTraversalStack - kSmemStack;    // deeper levels spill to local memory (never on the bench scenes)

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

// Rays and hits stream through once per launch while the BVH is re-read by every ray: their loads and stores carry the streaming
// (evict-first) cache operator so that they do not push nodes and triangles out of L2 (round-1 capture: 7.5 GB of DRAM traffic per
// launch against 1 GB compulsory).  -DLCB_NO_STREAM_HINTS restores plain accesses for A/B runs.
#ifdef LCB_NO_STREAM_HINTS
#define LCB_LD_STREAM(P) __ldg(P)
#define LCB_ST_STREAM(P, V) (*(P) = (V))
#else
#define LCB_LD_STREAM(P) __ldcs(P)
#define LCB_ST_STREAM(P, V) __stcs(P, V)
#endif

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
    float tmin = 0.f, tbest = 0.f;  // tbest starts as the ray's tmax and only shrinks: also the upper end of the interval triangles are tested against
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
                        const float4 ra = LCB_LD_STREAM(rays + 2 * ray_idx), rb = LCB_LD_STREAM(rays + 2 * ray_idx + 1);
                        setup_world(r, ra, rb);
                        tmin = ra.w; tbest = rb.w;
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
#ifdef LCB_COUNT_HOT
            if (COUNTERS && cur_inst != kNone && (uint32_t)(node - nodes) < (uint32_t)LCB_COUNT_HOT) n_inst++;
#endif
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
                    if (canonical_triangle(r, tmin, tbest, v0, v1, v2, t, V, W, det)) {
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
                    // procedural instances (flags bit 2) only produce candidates for a RayQuery's on_procedural_hit callback, which the
                    // batch entry points do not have: they are skipped, like user geometry without an intersect function in Embree
                    if ((meta.x & mask) != 0u && (meta.z & 4u) == 0u) {
                        if (Gt.y) LCB_PUSH(Gt)
                        if (G.y & 0xff000000u) LCB_PUSH(G)
                        const uint4 ptrs = __ldg(reinterpret_cast<const uint4 *>(rec) + 3);
                        nodes = reinterpret_cast<const WideNode *>(((unsigned long long)ptrs.y << 32) | ptrs.x);
                        tris = reinterpret_cast<const PackedTri *>(((unsigned long long)ptrs.w << 32) | ptrs.z);
                        LCB_PUSH(make_uint2(enter_instance(r, meta.z, rec) ? 1u : 0u, 0u))  // sentinel: below it lies world space (x = 1: same ray setup)
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
            if (Gt.y != 0u && (G.y & 0xff000000u) != 0u && sp < kPostponeLimit) { LCB_PUSH(Gt) Gt.y = 0u; }
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
                if (e.x == 0u) {
                    const float4 ra = __ldg(rays + 2 * ray_idx), rb = __ldg(rays + 2 * ray_idx + 1);
                    setup_world(r, ra, rb);
                }
            }
            if (retire) {
                if (ANY) {
                    LCB_ST_STREAM(reinterpret_cast<uint32_t *>(out) + ray_idx, hit_inst != kNone ? 1u : 0u);
                } else {
                    uint2 *o = reinterpret_cast<uint2 *>(out) + 3 * ray_idx;
                    LCB_ST_STREAM(o, make_uint2(hit_inst, hit_prim));
                    // barycentrics are formed by k_refine; the pad word carries the winning PackedTri slot to it
                    LCB_ST_STREAM(o + 2, make_uint2(__float_as_uint(tbest), hit_slot));  // a miss still holds the ray's tmax
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

// Accels that hold curve instances: one ray per thread through trace_one_impl<…, CURVES = true> (the persistent kernel above keeps
// its register budget for triangles).  Same records as k_trace + k_refine write.
struct FilterHook {
    CandidateFilter f;
    __device__ __forceinline__ int triangle(uint32_t inst, uint32_t prim, float u, float v, float) const { return candidate_commits(f, inst, prim, u, v) ? 1 : 0; }
    __device__ __forceinline__ int procedural(uint32_t, uint32_t, float, float &) const { return 0; }
};
template <int MODE>
__global__ void __launch_bounds__(128) k_trace_curves(AccelView acc, const float4 *__restrict__ rays, void *__restrict__ out, unsigned long long count, uint32_t mask,
                                                      CandidateFilter filter) {
    constexpr bool ANY = MODE == kAny;
    constexpr bool QUERY = MODE == kQueryAll || MODE == kQueryAny;
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const float4 ra = __ldg(rays + 2 * i), rb = __ldg(rays + 2 * i + 1);
    FilterHook hook{filter};
    const DeviceHit h = trace_one_impl<ANY, QUERY, FilterHook, true>(acc, ra, rb, mask, MODE == kQueryAny, hook);
    if (ANY) { reinterpret_cast<uint32_t *>(out)[i] = h.inst != kNone ? 1u : 0u; return; }
    uint2 *o = reinterpret_cast<uint2 *>(out) + 3 * i;
    o[0] = make_uint2(h.inst, h.prim);
    o[1] = make_uint2(__float_as_uint(h.u), __float_as_uint(h.v));
    if (QUERY) o[2] = h.inst != kNone ? make_uint2(1u, __float_as_uint(h.t)) : make_uint2(0u, 0u);
    else o[2] = make_uint2(__float_as_uint(h.t), 0u);
}

// Second pass of closest-hit queries: barycentrics of every hit, one thread per ray, fully converged.  The f64
// evaluation is the reported value (oracle.c refine_bary); if its determinant vanishes the canonical fp32 ones stand.
// COMMITTED selects the record layout: SurfaceHit {inst, prim, u, v, t, pad} or CommittedHit {inst, prim, u, v, hit_type, t}.
template <bool COMMITTED>
__global__ void __launch_bounds__(256) k_refine(AccelView acc, const float4 *__restrict__ rays, uint2 *__restrict__ hits, unsigned long long count) {
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const uint2 h0 = LCB_LD_STREAM(hits + 3 * i), h2 = LCB_LD_STREAM(hits + 3 * i + 2);
    if (h0.x == kNone) {
        LCB_ST_STREAM(hits + 3 * i + 1, make_uint2(0u, 0u));
        // a RayQuery that commits nothing leaves its zero-initialised CommittedHit (cpu_resource.h:320-331): hit_type Miss, t = 0
        LCB_ST_STREAM(hits + 3 * i + 2, COMMITTED ? make_uint2(0u /* HitType::Miss */, 0u) : make_uint2(h2.x, 0u));
        return;
    }
    const float4 *rec = reinterpret_cast<const float4 *>(acc.instances + h0.x);
    const float4 m0 = __ldg(rec), m1 = __ldg(rec + 1), m2 = __ldg(rec + 2);
    const uint4 ptrs = __ldg(reinterpret_cast<const uint4 *>(rec) + 3);
    const float4 *tp = reinterpret_cast<const float4 *>(reinterpret_cast<const PackedTri *>(((unsigned long long)ptrs.w << 32) | ptrs.z) + h2.y);
    const float4 ra = LCB_LD_STREAM(rays + 2 * i), rb = LCB_LD_STREAM(rays + 2 * i + 1);
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
    LCB_ST_STREAM(hits + 3 * i + 1, make_uint2(__float_as_uint(u), __float_as_uint(v)));
    LCB_ST_STREAM(hits + 3 * i + 2, COMMITTED ? make_uint2(1u /* HitType::Triangle */, h2.x) : make_uint2(h2.x, 0u));
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
            TraceCounters *ctr, LaunchCounter &lc, const CandidateFilter &filter = CandidateFilter{0, 0.f, nullptr, nullptr}, unsigned grid_limit = 0) {
    constexpr bool ANY = MODE == kAny;
    if (a.flags & 1u) {  // curve instances present
        if (count == 0) return;
        k_trace_curves<MODE><<<(unsigned)((count + 127) / 128), 128, 0, s>>>(a, reinterpret_cast<const float4 *>(rays), out, count, mask, filter);
        lc.count++;
        return;
    }
    // resident grid of this instantiation: a function-local static's initialiser runs once even when several host threads dispatch
    struct Resident { int blocks_per_sm = 0, sms = 0; };
    static const Resident res = [] {
        Resident r; int dev = 0; cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&r.sms, cudaDevAttrMultiProcessorCount, dev);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&r.blocks_per_sm, k_trace<MODE, COUNTERS>, kTraceThreads, 0);
        if (r.blocks_per_sm < 1) r.blocks_per_sm = 1;
        return r;
    }();
    const int blocks_per_sm = res.blocks_per_sm, sms = res.sms;
    unsigned long long want = (count + kTraceThreads - 1) / kTraceThreads;
    unsigned long long grid = (unsigned long long)sms * blocks_per_sm;
    if (grid > want) grid = want;
    if (grid_limit && grid > grid_limit) grid = grid_limit;  // chunked host pipeline: leave CTA slots to the neighbouring chunk's launch
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
            const bool in_alt = sort_pairs(s, n, keys, vals, keys_alt, vals_alt, p + 2 * kb + 2 * vb, 0, passes, 8, lc);
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
                   TraceCounters *counters, LaunchCounter &lc, unsigned grid_limit) {
    const CandidateFilter none{0, 0.f, nullptr, nullptr};
    if (counters) launch<kClosest, true>(s, a, rays, hits, count, mask, work_counter, counters, lc, none, grid_limit);
    else launch<kClosest, false>(s, a, rays, hits, count, mask, work_counter, nullptr, lc, none, grid_limit);
}

void trace_any(cudaStream_t s, const AccelView &a, const void *rays, uint32_t *occluded, uint64_t count, uint32_t mask, unsigned long long *work_counter,
               LaunchCounter &lc, unsigned grid_limit) {
    const CandidateFilter none{0, 0.f, nullptr, nullptr};
    launch<kAny, false>(s, a, rays, occluded, count, mask, work_counter, nullptr, lc, none, grid_limit);
}

void ray_query(cudaStream_t s, const AccelView &a, const void *rays, void *committed_hits, uint64_t count, uint32_t mask, bool terminate_on_first,
               const CandidateFilter &filter, unsigned long long *work_counter, LaunchCounter &lc) {
    if (terminate_on_first) launch<kQueryAny, false>(s, a, rays, committed_hits, count, mask, work_counter, nullptr, lc, filter);
    else launch<kQueryAll, false>(s, a, rays, committed_hits, count, mask, work_counter, nullptr, lc, filter);
}

}  // namespace lcb
