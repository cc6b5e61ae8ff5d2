// bvh_build.cu — LBVH builder for sm_100a: primitive boxes + centroid bounds, 48-bit Morton
// codes, onesweep sort (radix_sort.cu), fused bottom-up hierarchy emission + AABB refit with
// atomic arrival flags (after Apetrei 2014), and collapse of the binary tree into 8-wide
// 128-byte quantised nodes with packed 64-byte leaf triangles.
//
// Replaces what the reference delegates to rtcCommitScene (cpu/accel.rs:258,439) /
// optixAccelBuild (cuda_primitive.cpp:57-60).  No reference source exists for any of it.
#include "build.cuh"
#include "trace_device.cuh"
#include <cfloat>
#include <cstdlib>

namespace lcb {

namespace {

constexpr uint32_t kLeafBit = 0x80000000u;
#ifndef LCB_LEAF_MAX
#define LCB_LEAF_MAX 2
#endif
constexpr int kLeafMax = LCB_LEAF_MAX;  // primitives per leaf child (unary count in 3 bits)

__device__ __forceinline__ float fmin3(float a, float b, float c) { return fminf(a, fminf(b, c)); }
__device__ __forceinline__ float fmax3(float a, float b, float c) { return fmaxf(a, fmaxf(b, c)); }

__global__ void k_init_header(BuildHeader *h, int *flags, uint32_t n_flags) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) {
        for (int k = 0; k < 3; k++) { h->bounds_lo[k] = 0x7fffffff; h->bounds_hi[k] = (int)0x80000000; h->root_lo[k] = 0.f; h->root_hi[k] = 0.f; }
        h->root = 0; h->node_count = 1; h->prim_count = 0; h->emitted = 0; h->bar_count = 0; h->bar_release = 0; h->max_depth = 0; h->error = 0;
        h->prim_area_sum = 0.f; h->pad2 = 0.f;
    }
    if (i < 48) h->level_end[i] = 0;  // the host reads the header back whole (builder choice, compaction)
    for (uint32_t j = i; j < n_flags; j += gridDim.x * blockDim.x) flags[j] = -1;
}

// block-reduce the centroid (shuffles, then one shared-memory round) and fold it into the header's ordered-int
// bounds: 6 atomics per block instead of per warp
__device__ __forceinline__ void reduce_centroid_bounds(float c[3], bool valid, BuildHeader *h, float area = 0.f) {
    __shared__ float s_lo[8][3], s_hi[8][3], s_area[8];
    float lo[3], hi[3];
    if (!valid) area = 0.f;
#pragma unroll
    for (int k = 0; k < 3; k++) { lo[k] = valid ? c[k] : FLT_MAX; hi[k] = valid ? c[k] : -FLT_MAX; }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
#pragma unroll
        for (int k = 0; k < 3; k++) {
            lo[k] = fminf(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], off));
            hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], off));
        }
        area += __shfl_xor_sync(0xffffffffu, area, off);
    }
    const int warp = threadIdx.x >> 5, n_warps = (blockDim.x + 31) >> 5;
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int k = 0; k < 3; k++) { s_lo[warp][k] = lo[k]; s_hi[warp][k] = hi[k]; }
        s_area[warp] = area;
    }
    __syncthreads();
    if (threadIdx.x == 3) {
        float a = 0.f;
        for (int w = 0; w < n_warps; w++) a += s_area[w];
        if (a > 0.f) atomicAdd(&h->prim_area_sum, a);
    }
    if (threadIdx.x < 3) {
        const int k = threadIdx.x;
        float l = s_lo[0][k], u = s_hi[0][k];
        for (int w = 1; w < n_warps; w++) { l = fminf(l, s_lo[w][k]); u = fmaxf(u, s_hi[w][k]); }
        if (l <= u) {
            atomicMin(&h->bounds_lo[k], float_to_ordered(l));
            atomicMax(&h->bounds_hi[k], float_to_ordered(u));
        }
    }
}

__device__ __forceinline__ void load_triangle(const TriangleInput &in, uint32_t prim, float a[3], float b[3], float c[3]) {
    const uint32_t *ix = reinterpret_cast<const uint32_t *>(in.indices + (size_t)prim * 12);
    const uint32_t i0 = ix[0], i1 = ix[1], i2 = ix[2];
    const float *pa = reinterpret_cast<const float *>(in.vertices + (size_t)i0 * in.vertex_stride);
    const float *pb = reinterpret_cast<const float *>(in.vertices + (size_t)i1 * in.vertex_stride);
    const float *pc = reinterpret_cast<const float *>(in.vertices + (size_t)i2 * in.vertex_stride);
#pragma unroll
    for (int k = 0; k < 3; k++) { a[k] = pa[k]; b[k] = pb[k]; c[k] = pc[k]; }
}

__global__ void __launch_bounds__(256) k_triangle_boxes(TriangleInput in, uint32_t n, PrimBox *boxes, BuildHeader *h) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    float cen[3] = {0, 0, 0};
    float area = 0.f;
    bool valid = i < n;
    if (valid) {
        float a[3], b[3], c[3];
        load_triangle(in, i, a, b, c);
        PrimBox pb;
#pragma unroll
        for (int k = 0; k < 3; k++) {
            pb.lo[k] = fmin3(a[k], b[k], c[k]);
            pb.hi[k] = fmax3(a[k], b[k], c[k]);
            cen[k] = 0.5f * pb.lo[k] + 0.5f * pb.hi[k];
        }
        pb.pad0 = pb.pad1 = 0;
        reinterpret_cast<float4 *>(boxes)[2 * (size_t)i] = make_float4(pb.lo[0], pb.lo[1], pb.lo[2], 0.f);
        reinterpret_cast<float4 *>(boxes)[2 * (size_t)i + 1] = make_float4(pb.hi[0], pb.hi[1], pb.hi[2], 0.f);
        const float dx = pb.hi[0] - pb.lo[0], dy = pb.hi[1] - pb.lo[1], dz = pb.hi[2] - pb.lo[2];
        area = dx * dy + dy * dz + dz * dx;
    }
    reduce_centroid_bounds(cen, valid, h, area);
}

// Procedural primitives (ProceduralPrimitiveBuild, api_types:633-641): the user's AABBs {min[3], max[3]} (rtx.rs:339-345, 24 B) are
// the primitive boxes (GeometryImpl::build_procedural's bounds_func, cpu/accel.rs:93-104).
__global__ void __launch_bounds__(256) k_aabb_boxes(const uint8_t *__restrict__ aabbs, uint32_t n, PrimBox *boxes, BuildHeader *h) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    float cen[3] = {0, 0, 0};
    float area = 0.f;
    bool valid = i < n;
    if (valid) {
        const float *a = reinterpret_cast<const float *>(aabbs + (size_t)i * 24);
        float lo[3], hi[3];
#pragma unroll
        for (int k = 0; k < 3; k++) { lo[k] = fminf(a[k], a[3 + k]); hi[k] = fmaxf(a[k], a[3 + k]); cen[k] = 0.5f * lo[k] + 0.5f * hi[k]; }
        reinterpret_cast<float4 *>(boxes)[2 * (size_t)i] = make_float4(lo[0], lo[1], lo[2], 0.f);
        reinterpret_cast<float4 *>(boxes)[2 * (size_t)i + 1] = make_float4(hi[0], hi[1], hi[2], 0.f);
        const float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
        area = dx * dy + dy * dz + dz * dx;
    }
    reduce_centroid_bounds(cen, valid, h, area);
}

// leaf records of a procedural BLAS: the primitive's box and id in the PackedTri slot the collapse assigned
__global__ void __launch_bounds__(256) k_pack_aabbs(const uint8_t *__restrict__ aabbs, PackedTri *tris, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t prim = tris[i].prim;
    const float *a = reinterpret_cast<const float *>(aabbs + (size_t)prim * 24);
    float4 *o = reinterpret_cast<float4 *>(&tris[i]);
    o[0] = make_float4(fminf(a[0], a[3]), fminf(a[1], a[4]), fminf(a[2], a[5]), __uint_as_float(prim));
    o[1] = make_float4(fmaxf(a[0], a[3]), fmaxf(a[1], a[4]), fmaxf(a[2], a[5]), 0.f);
    o[2] = make_float4(0.f, 0.f, 0.f, 0.f);
}

// Curve pieces (CurveBuild): box of the two end spheres, rounded outward.
__global__ void __launch_bounds__(256) k_curve_boxes(CurveInput in, uint32_t n, PrimBox *boxes, BuildHeader *h) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    float cen[3] = {0, 0, 0};
    float area = 0.f;
    bool valid = i < n;
    if (valid) {
        float4 A, B;
        curve_piece(in.cps, in.cp_stride, in.segs, in.basis, i / in.pieces, i % in.pieces, A, B);
        const float ra = fabsf(A.w), rb = fabsf(B.w);
        const float pa[3] = {A.x, A.y, A.z}, pb[3] = {B.x, B.y, B.z};
        float lo[3], hi[3];
#pragma unroll
        // the canonical cone test decides in fp32: pad by a fraction of the radius and of the piece's extent so that a ray it
        // accepts never misses the box
        const float pad = fmaxf(ra, rb) * (1.0f / 256.0f) + fmaxf(fmaxf(fabsf(pb[0] - pa[0]), fabsf(pb[1] - pa[1])), fabsf(pb[2] - pa[2])) * (1.0f / 4096.0f);
        for (int k = 0; k < 3; k++) {
            lo[k] = __fsub_rd(fminf(__fsub_rd(pa[k], ra), __fsub_rd(pb[k], rb)), pad);
            hi[k] = __fadd_ru(fmaxf(__fadd_ru(pa[k], ra), __fadd_ru(pb[k], rb)), pad);
            cen[k] = 0.5f * lo[k] + 0.5f * hi[k];
        }
        reinterpret_cast<float4 *>(boxes)[2 * (size_t)i] = make_float4(lo[0], lo[1], lo[2], 0.f);
        reinterpret_cast<float4 *>(boxes)[2 * (size_t)i + 1] = make_float4(hi[0], hi[1], hi[2], 0.f);
        const float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
        area = dx * dy + dy * dz + dz * dx;
    }
    reduce_centroid_bounds(cen, valid, h, area);
}

// leaf records of a curve BLAS: the piece's two spheres, its segment and its parameter range (CurveSeg)
__global__ void __launch_bounds__(256) k_pack_curves(CurveInput in, PackedTri *slots, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t j = slots[i].prim, seg = j / in.pieces, k = j % in.pieces;
    float4 A, B;
    curve_piece(in.cps, in.cp_stride, in.segs, in.basis, seg, k, A, B);
    const float du = 1.0f / (float)in.pieces;
    float4 *o = reinterpret_cast<float4 *>(&slots[i]);
    o[0] = make_float4(A.x, A.y, A.z, __uint_as_float(seg));
    o[1] = make_float4(B.x, B.y, B.z, A.w);
    o[2] = make_float4(B.w, (float)k * du, du, __uint_as_float(n + seg));  // coefficient slot of the segment (k_curve_coefs)
}

// power-basis coefficients of every cubic segment, one 64-byte slot each, behind the n leaf slots (read by refine_curve_hit)
__global__ void __launch_bounds__(256) k_curve_coefs(CurveInput in, PackedTri *slots, uint32_t n, uint32_t n_segs) {
    const uint32_t seg = blockIdx.x * blockDim.x + threadIdx.x;
    if (seg >= n_segs) return;
    float4 a[4];
    curve_coefficients(in.cps, in.cp_stride, in.segs, in.basis, seg, a);
    float4 *o = reinterpret_cast<float4 *>(&slots[n + seg]);
    o[0] = a[0]; o[1] = a[1]; o[2] = a[2]; o[3] = a[3];
}

// World-space box of an instance: union of the BLAS root's (conservatively decoded) child
// boxes, 8 corners through the affine, padded for the fp32 mismatch between M and M^-1.
__global__ void __launch_bounds__(128) k_instance_boxes(const uint32_t *active, uint32_t n, const InstanceRec *insts, PrimBox *boxes, BuildHeader *h) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    float cen[3] = {0, 0, 0};
    bool valid = i < n;
    if (valid) {
        const InstanceRec &rec = insts[active[i]];
        const WideNode &root = rec.nodes[0];
        float olo[3], ohi[3];
#pragma unroll
        for (int k = 0; k < 3; k++) {
            float scale = __uint_as_float((uint32_t)root.e[k] << 23);
            uint32_t qmin = kQMax, qmax = 0;
            for (int c = 0; c < 8; c++) {
                if (root.meta[c] == 0) continue;
                qmin = min(qmin, (uint32_t)root.q[k][0][c]);
                qmax = max(qmax, (uint32_t)root.q[k][1][c]);
            }
            olo[k] = __fmaf_rd((float)qmin, scale, root.org[k]);
            ohi[k] = __fmaf_ru((float)qmax, scale, root.org[k]);
        }
        float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
        for (int corner = 0; corner < 8; corner++) {
            float p[3] = {(corner & 1) ? ohi[0] : olo[0], (corner & 2) ? ohi[1] : olo[1], (corner & 4) ? ohi[2] : olo[2]};
#pragma unroll
            for (int r = 0; r < 3; r++) {
                const float *m = rec.affine + 4 * r;
                float w = fmaf(m[0], p[0], fmaf(m[1], p[1], fmaf(m[2], p[2], m[3])));
                lo[r] = fminf(lo[r], w); hi[r] = fmaxf(hi[r], w);
            }
        }
        float mag = 0.f;
#pragma unroll
        for (int k = 0; k < 3; k++) mag = fmaxf(mag, fmaxf(fmaxf(fabsf(lo[k]), fabsf(hi[k])), hi[k] - lo[k]));
        const float pad = mag * (1.0f / 16384.0f) + FLT_MIN;
#pragma unroll
        for (int k = 0; k < 3; k++) { lo[k] -= pad; hi[k] += pad; cen[k] = 0.5f * lo[k] + 0.5f * hi[k]; }
        reinterpret_cast<float4 *>(boxes)[2 * (size_t)i] = make_float4(lo[0], lo[1], lo[2], 0.f);
        reinterpret_cast<float4 *>(boxes)[2 * (size_t)i + 1] = make_float4(hi[0], hi[1], hi[2], 0.f);
    }
    reduce_centroid_bounds(cen, valid, h);
}

__device__ __forceinline__ uint64_t expand21(uint32_t x) {
    uint64_t v = x & 0x1fffffull;
    v = (v | v << 32) & 0x1f00000000ffffull;
    v = (v | v << 16) & 0x1f0000ff0000ffull;
    v = (v | v << 8) & 0x100f00f00f00f00full;
    v = (v | v << 4) & 0x10c30c30c30c30c3ull;
    v = (v | v << 2) & 0x1249249249249249ull;
    return v;
}

__global__ void __launch_bounds__(256) k_morton(const PrimBox *__restrict__ boxes, uint32_t n, const BuildHeader *__restrict__ h,
                                                uint64_t *__restrict__ keys, uint32_t *__restrict__ vals, int drop_bits) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 lo = reinterpret_cast<const float4 *>(boxes)[2 * (size_t)i];
    float4 hi = reinterpret_cast<const float4 *>(boxes)[2 * (size_t)i + 1];
    float c[3] = {0.5f * lo.x + 0.5f * hi.x, 0.5f * lo.y + 0.5f * hi.y, 0.5f * lo.z + 0.5f * hi.z};
    uint32_t q[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        float blo = ordered_to_float(h->bounds_lo[k]), bhi = ordered_to_float(h->bounds_hi[k]);
        float ext = bhi - blo;
        float t = ext > 0.f ? (c[k] - blo) / ext : 0.f;
        t = fminf(fmaxf(t, 0.f), 1.f);
        q[k] = min((uint32_t)(t * 2097152.0f), 2097151u);
    }
    // the top 8 * passes bits of the 63-bit code, one 8-bit sort pass each; equal keys are split by position
    keys[i] = (expand21(q[0]) | (expand21(q[1]) << 1) | (expand21(q[2]) << 2)) >> drop_bits;
    vals[i] = i;
}

// ---- fused hierarchy + refit ---------------------------------------------------------------
// Internal node i separates sorted leaves i and i+1.  delta(i) orders the splits: the XOR of
// adjacent keys, with runs of equal keys split by position bits.
__device__ __forceinline__ uint64_t split_delta(const uint64_t *__restrict__ keys, uint32_t i) {
    uint64_t x = keys[i] ^ keys[i + 1];
    return x ? (x | (1ull << 63)) : (uint64_t)(i ^ (i + 1));
}

// Two phases.  (1) Warp-local: the 32 consecutive sorted leaves of a warp are merged with shuffles only — every lane owns the
// cluster that starts at its leaf; per step a cluster whose parent lies to its right merges with the next active cluster when
// that one's parent lies to its left (then both are children of internal node `right`); a cluster whose sibling is outside the
// warp's span leaves for phase 2.  31 of 32 internal nodes are emitted here without atomics, fences or L2 round trips.
// (2) Global: the classic bottom-up climb with one atomic exchange per arrival (Apetrei 2014): the first child to arrive at an
// internal node leaves its box there and stops, the second merges and continues.
__global__ void __launch_bounds__(128) k_hierarchy(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ prim, const PrimBox *__restrict__ boxes,
                                                   uint32_t n, BinNode *bin, int *flags, BuildHeader *h) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t lane = threadIdx.x & 31u, warp_base = i - lane;
    if (warp_base >= n) return;  // whole warp out of range
    const bool in_range = i < n;
    float4 lo = make_float4(0.f, 0.f, 0.f, 0.f), hi = lo;
    if (in_range) {
        const uint32_t p = prim[i];
        lo = reinterpret_cast<const float4 *>(boxes)[2 * (size_t)p];
        hi = reinterpret_cast<const float4 *>(boxes)[2 * (size_t)p + 1];
    }
    uint32_t left = i, right = i, cur = i | kLeafBit;
    if (n == 1) {
        if (i == 0) {
            h->root = cur;
            h->root_lo[0] = lo.x; h->root_lo[1] = lo.y; h->root_lo[2] = lo.z;
            h->root_hi[0] = hi.x; h->root_hi[1] = hi.y; h->root_hi[2] = hi.z;
        }
        return;
    }
    // delta of the split after leaf i (between i and i + 1), and of the split before the warp's first leaf
    const uint64_t d_mine = (in_range && i + 1 < n) ? split_delta(keys, i) : ~0ull;
    const uint64_t d_before = warp_base > 0 ? split_delta(keys, warp_base - 1) : ~0ull;
    bool local = in_range;    // still owned by phase 1
    bool climbing = false;    // left phase 1 for phase 2
    bool done = false;
    for (;;) {
        const uint32_t active = __ballot_sync(0xffffffffu, local);
        if (active == 0u) break;
        // my decision: is my parent the split to my right?
        const uint64_t d_right = __shfl_sync(0xffffffffu, d_mine, (right - warp_base) & 31u);
        const uint64_t d_left_in = __shfl_sync(0xffffffffu, d_mine, (left - 1u - warp_base) & 31u);
        const uint64_t d_left = left == warp_base ? d_before : d_left_in;
        const bool want_right = local && ((left == 0) || (right != n - 1 && d_right < d_left));
        const bool want_left = local && !want_right;
        const uint32_t next = (active & ~((2u << lane) - 1u)) ? (uint32_t)__ffs(active & ~((2u << lane) - 1u)) - 1u : 32u;
        const uint32_t next_wants_left = __ballot_sync(0xffffffffu, want_left);
        const bool merge = want_right && next < 32u && (next_wants_left >> next & 1u);
        const uint32_t merge_ballot = __ballot_sync(0xffffffffu, merge);
        // the cluster right before me merges with me this step: I am absorbed
        const uint32_t prev = (active & ((1u << lane) - 1u)) ? 31u - (uint32_t)__clz(active & ((1u << lane) - 1u)) : 32u;
        const bool absorbed = local && prev < 32u && (merge_ballot >> prev & 1u);
        // sibling data travels from lane `next` to me
        const uint32_t src = next & 31u;
        const float4 slo = make_float4(__shfl_sync(0xffffffffu, lo.x, src), __shfl_sync(0xffffffffu, lo.y, src), __shfl_sync(0xffffffffu, lo.z, src), 0.f);
        const float4 shi = make_float4(__shfl_sync(0xffffffffu, hi.x, src), __shfl_sync(0xffffffffu, hi.y, src), __shfl_sync(0xffffffffu, hi.z, src), 0.f);
        const uint32_t s_cur = __shfl_sync(0xffffffffu, cur, src), s_right = __shfl_sync(0xffffffffu, right, src);
        if (merge) {
            const uint32_t parent = right;
            float4 *pn = reinterpret_cast<float4 *>(&bin[parent]);
            pn[0] = make_float4(lo.x, lo.y, lo.z, __uint_as_float(cur));
            pn[1] = make_float4(hi.x, hi.y, hi.z, __uint_as_float(right - left + 1u));
            pn[2] = make_float4(slo.x, slo.y, slo.z, __uint_as_float(s_cur));
            pn[3] = make_float4(shi.x, shi.y, shi.z, __uint_as_float(s_right - right));
            lo.x = fminf(lo.x, slo.x); lo.y = fminf(lo.y, slo.y); lo.z = fminf(lo.z, slo.z);
            hi.x = fmaxf(hi.x, shi.x); hi.y = fmaxf(hi.y, shi.y); hi.z = fmaxf(hi.z, shi.z);
            right = s_right; cur = parent;
            if (left == 0 && right == n - 1) {
                h->root = parent;
                h->root_lo[0] = lo.x; h->root_lo[1] = lo.y; h->root_lo[2] = lo.z;
                h->root_hi[0] = hi.x; h->root_hi[1] = hi.y; h->root_hi[2] = hi.z;
                local = false; done = true;
            }
        } else if (absorbed) {
            local = false; done = true;
        } else if ((want_right && next == 32u) || (want_left && prev == 32u)) {
            // no phase-1 cluster on the side my sibling will come from: it lives in (or will be formed by clusters of) another warp.
            // Only the outermost clusters can leave, so the remaining ones stay contiguous.
            local = false; climbing = true;
        }
    }
    if (!climbing || done) return;
    while (true) {
        const uint32_t count = right - left + 1;
        const bool parent_on_right = (left == 0) || (right != n - 1 && split_delta(keys, right) < split_delta(keys, left - 1));
        uint32_t parent;
        float4 slo, shi;
        if (parent_on_right) {  // we are the left child of internal node `right`
            parent = right;
            float4 *pn = reinterpret_cast<float4 *>(&bin[parent]);
            pn[0] = make_float4(lo.x, lo.y, lo.z, __uint_as_float(cur));
            pn[1] = make_float4(hi.x, hi.y, hi.z, __uint_as_float(count));
            __threadfence();
            int other = atomicExch(&flags[parent], (int)left);
            if (other == -1) return;
            right = (uint32_t)other;  // the sibling fenced before its exchange; its box is read at L2 (__ldcg) below
            slo = __ldcg(pn + 2); shi = __ldcg(pn + 3);
        } else {  // right child of internal node `left - 1`
            parent = left - 1;
            float4 *pn = reinterpret_cast<float4 *>(&bin[parent]);
            pn[2] = make_float4(lo.x, lo.y, lo.z, __uint_as_float(cur));
            pn[3] = make_float4(hi.x, hi.y, hi.z, __uint_as_float(count));
            __threadfence();
            int other = atomicExch(&flags[parent], (int)right);
            if (other == -1) return;
            left = (uint32_t)other;
            slo = __ldcg(pn); shi = __ldcg(pn + 1);
        }
        lo.x = fminf(lo.x, slo.x); lo.y = fminf(lo.y, slo.y); lo.z = fminf(lo.z, slo.z);
        hi.x = fmaxf(hi.x, shi.x); hi.y = fmaxf(hi.y, shi.y); hi.z = fmaxf(hi.z, shi.z);
        cur = parent;
        if (left == 0 && right == n - 1) {
            h->root = parent;
            h->root_lo[0] = lo.x; h->root_lo[1] = lo.y; h->root_lo[2] = lo.z;
            h->root_hi[0] = hi.x; h->root_hi[1] = hi.y; h->root_hi[2] = hi.z;
            return;
        }
    }
}

// ---- PLOC: parallel locally-ordered clustering (Meister & Bittner 2018; fused search + merge after Benthin et al. 2022) ----
// The alternative to k_hierarchy for AccelUsageHint::FastTrace.  The Morton-sorted primitives start as one cluster each.  Every
// iteration, each cluster looks r positions to the left and right for the neighbour whose union box has the smallest surface
// area; mutual nearest neighbours merge into a new binary node, and the surviving clusters are compacted in order.  The result
// is a BinNode array of the same shape k_hierarchy produces (child boxes, ids, leaf counts), so the collapse is shared.
// Per iteration: k_ploc_step (search + merge + survivor flags + per-block counts), k_ploc_scan (one block: exclusive scan of the
// block counts, publishes the new cluster count), k_ploc_compact (ordered scatter).  A cluster is two float4:
// (lo.xyz, id) and (hi.xyz, leaf count).
constexpr int kPlocTile = 256;
constexpr int kPlocMaxRadius = 16;


__global__ void __launch_bounds__(256) k_ploc_init(const uint32_t *__restrict__ prim, const PrimBox *__restrict__ boxes, uint32_t n, float4 *clusters, PlocState *st) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) { st->n_clusters = n; st->n_nodes = 0; st->iterations = 0; }
    if (i >= n) return;
    const uint32_t p = prim[i];
    const float4 lo = reinterpret_cast<const float4 *>(boxes)[2 * (size_t)p], hi = reinterpret_cast<const float4 *>(boxes)[2 * (size_t)p + 1];
    clusters[2 * (size_t)i] = make_float4(lo.x, lo.y, lo.z, __uint_as_float(i | kLeafBit));
    clusters[2 * (size_t)i + 1] = make_float4(hi.x, hi.y, hi.z, __uint_as_float(1u));
}

__device__ __forceinline__ float union_half_area(const float4 alo, const float4 ahi, const float4 blo, const float4 bhi) {
    const float dx = fmaxf(ahi.x, bhi.x) - fminf(alo.x, blo.x), dy = fmaxf(ahi.y, bhi.y) - fminf(alo.y, blo.y), dz = fmaxf(ahi.z, bhi.z) - fminf(alo.z, blo.z);
    return dx * dy + dy * dz + dz * dx;
}

__global__ void __launch_bounds__(kPlocTile) k_ploc_step(const float4 *__restrict__ in, float4 *__restrict__ out, uint32_t *__restrict__ keep, uint32_t *__restrict__ block_counts,
                                                         BinNode *bin, PlocState *st, int radius, int forced) {
    __shared__ float4 s_lo[kPlocTile + 4 * kPlocMaxRadius], s_hi[kPlocTile + 4 * kPlocMaxRadius];
    __shared__ int s_nn[kPlocTile + 2 * kPlocMaxRadius];
    __shared__ uint32_t s_warp[kPlocTile / 32], s_node_base;
    const uint32_t n = st->n_clusters;
    if (n <= 1) return;
    // persistent tiles: the grid is sized once per build, the live cluster count is read on the device
    for (int base = (int)(blockIdx.x * kPlocTile); (uint32_t)base < n; base += (int)(gridDim.x * kPlocTile)) {
    const int r = radius, span = kPlocTile + 4 * r, first = base - 2 * r;
    const uint32_t tile_index = (uint32_t)base / kPlocTile;
    for (int t = threadIdx.x; t < span; t += kPlocTile) {
        const int i = first + t;
        if (i >= 0 && (uint32_t)i < n) { s_lo[t] = in[2 * (size_t)i]; s_hi[t] = in[2 * (size_t)i + 1]; }
    }
    __syncthreads();
    // nearest neighbour (smallest union area, ties -> lower index) of every cluster of the tile and of r clusters on either side
    for (int t = threadIdx.x; t < kPlocTile + 2 * r; t += kPlocTile) {
        const int i = base - r + t;
        int best = -1;
        if (i >= 0 && (uint32_t)i < n) {
            const float4 lo = s_lo[i - first], hi = s_hi[i - first];
            if (forced) {  // pair (2k, 2k+1): halves the cluster count whatever the geometry (see run_pipeline_after_boxes)
                best = (uint32_t)(i ^ 1) < n ? (i ^ 1) : -1;
            } else {
                // smallest union area; among equal areas the "buddy" i ^ 1 first (so that a run of identical boxes pairs up
                // instead of forming a chain with a single mutual pair), then the lower index.  The order is symmetric in (i, j).
                float best_a = FLT_MAX;
                bool best_buddy = false;
                for (int j = i - r; j <= i + r; j++) {
                    if (j == i || j < 0 || (uint32_t)j >= n) continue;
                    const float a = union_half_area(lo, hi, s_lo[j - first], s_hi[j - first]);
                    const bool buddy = (i ^ 1) == j;
                    if (best < 0 || a < best_a || (a == best_a && buddy && !best_buddy)) { best_a = a; best = j; best_buddy = buddy; }
                }
            }
        }
        s_nn[t] = best;
    }
    __syncthreads();
    const int i = base + (int)threadIdx.x;
    bool valid = (uint32_t)i < n, merge = false, survive = valid;
    int j = -1;
    if (valid) {
        j = s_nn[i - (base - r)];
        const bool mutual = j >= 0 && s_nn[j - (base - r)] == i;
        merge = mutual && i < j;
        survive = !(mutual && i > j);
    }
    // node ids: one atomic per block
    const uint32_t merge_ballot = __ballot_sync(0xffffffffu, merge), lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    if (lane == 0) s_warp[warp] = __popc(merge_ballot);
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t tot = 0;
        for (int w = 0; w < kPlocTile / 32; w++) { const uint32_t c = s_warp[w]; s_warp[w] = tot; tot += c; }
        s_node_base = tot ? atomicAdd(&st->n_nodes, tot) : 0u;
    }
    __syncthreads();
    if (valid) {
        float4 lo = s_lo[i - first], hi = s_hi[i - first];
        if (merge) {
            const uint32_t id = s_node_base + s_warp[warp] + __popc(merge_ballot & ((1u << lane) - 1u));
            const float4 olo = s_lo[j - first], ohi = s_hi[j - first];
            float4 *pn = reinterpret_cast<float4 *>(&bin[id]);
            pn[0] = lo; pn[1] = hi; pn[2] = olo; pn[3] = ohi;   // (box, child id | leaf count) of the left and right child
            const uint32_t count = __float_as_uint(hi.w) + __float_as_uint(ohi.w);
            lo = make_float4(fminf(lo.x, olo.x), fminf(lo.y, olo.y), fminf(lo.z, olo.z), __uint_as_float(id));
            hi = make_float4(fmaxf(hi.x, ohi.x), fmaxf(hi.y, ohi.y), fmaxf(hi.z, ohi.z), __uint_as_float(count));
        }
        out[2 * (size_t)i] = lo; out[2 * (size_t)i + 1] = hi;
        keep[i] = survive ? 1u : 0u;
    }
    __syncthreads();
    const uint32_t keep_ballot = __ballot_sync(0xffffffffu, survive);
    if (lane == 0) s_warp[warp] = __popc(keep_ballot);
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t tot = 0;
        for (int w = 0; w < kPlocTile / 32; w++) tot += s_warp[w];
        block_counts[tile_index] = tot;
    }
    __syncthreads();  // shared arrays are reused by the next tile
    }
}

// one block: exclusive scan of the per-tile survivor counts, in place; publishes the new cluster count
__global__ void __launch_bounds__(1024) k_ploc_scan(uint32_t *block_counts, PlocState *st) {
    __shared__ uint32_t s_warp[32], s_carry;
    const uint32_t n = st->n_clusters;
    if (n <= 1) { if (threadIdx.x == 0) st->pad = 0; return; }  // nothing left to compact
    const uint32_t n_blocks = (n + kPlocTile - 1) / kPlocTile;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (uint32_t b0 = 0; b0 < n_blocks; b0 += 1024) {
        const uint32_t b = b0 + threadIdx.x;
        const uint32_t v = b < n_blocks ? block_counts[b] : 0u;
        uint32_t x = v;
        for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, x, o); if ((threadIdx.x & 31) >= (uint32_t)o) x += y; }
        if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = x;
        __syncthreads();
        if (threadIdx.x < 32) {
            uint32_t w = s_warp[threadIdx.x];
            for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, w, o); if (threadIdx.x >= (uint32_t)o) w += y; }
            s_warp[threadIdx.x] = w;
        }
        __syncthreads();
        const uint32_t incl = x + (threadIdx.x >= 32 ? s_warp[(threadIdx.x >> 5) - 1] : 0u) + s_carry;
        if (b < n_blocks) block_counts[b] = incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) { st->pad = n; st->n_clusters = s_carry; st->iterations++; }  // pad: the cluster count the compaction reads from
}

// ordered compaction of the survivors; `n_before` is implied by the block-count array (blocks past it hold stale data and exit)
__global__ void __launch_bounds__(kPlocTile) k_ploc_compact(const float4 *__restrict__ in, const uint32_t *__restrict__ keep, const uint32_t *__restrict__ block_offsets,
                                                            float4 *__restrict__ out, const PlocState *st) {
    __shared__ uint32_t s_warp[kPlocTile / 32];
    const uint32_t n_before = st->pad;
    if (n_before <= 1) return;
    for (uint32_t tile = blockIdx.x; tile * kPlocTile < n_before; tile += gridDim.x) {
    const uint32_t i = tile * kPlocTile + threadIdx.x;
    const bool k = i < n_before && keep[i] != 0u;
    const uint32_t ballot = __ballot_sync(0xffffffffu, k), lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    if (lane == 0) s_warp[warp] = __popc(ballot);
    __syncthreads();
    uint32_t off = block_offsets[tile];
    for (uint32_t w = 0; w < warp; w++) off += s_warp[w];
    if (k) {
        const uint32_t dst = off + __popc(ballot & ((1u << lane) - 1u));
        out[2 * (size_t)dst] = in[2 * (size_t)i]; out[2 * (size_t)dst + 1] = in[2 * (size_t)i + 1];
    }
    __syncthreads();
    }
}

__global__ void k_ploc_finish(const float4 *clusters, const PlocState *st, BuildHeader *h) {
    if (st->n_clusters != 1) { h->error = 3u; return; }
    const float4 lo = clusters[0], hi = clusters[1];
    h->root = __float_as_uint(lo.w);
    h->root_lo[0] = lo.x; h->root_lo[1] = lo.y; h->root_lo[2] = lo.z;
    h->root_hi[0] = hi.x; h->root_hi[1] = hi.y; h->root_hi[2] = hi.z;
}

// ---- collapse to 8-wide quantised nodes ----------------------------------------------------
struct Child {
    float lo[3], hi[3];
    uint32_t id;     // binary node id (kLeafBit => single sorted leaf)
    uint32_t count;  // leaves below
};

__device__ __forceinline__ float half_area(const Child &c) {
    float dx = c.hi[0] - c.lo[0], dy = c.hi[1] - c.lo[1], dz = c.hi[2] - c.lo[2];
    return dx * dy + dy * dz + dz * dx;
}

__device__ __forceinline__ void split_child(const BinNode *__restrict__ bin, const Child &c, Child &l, Child &r) {
    const float4 *pn = reinterpret_cast<const float4 *>(&bin[c.id]);
    float4 a = pn[0], b = pn[1], cc = pn[2], d = pn[3];
    l.lo[0] = a.x; l.lo[1] = a.y; l.lo[2] = a.z; l.id = __float_as_uint(a.w);
    l.hi[0] = b.x; l.hi[1] = b.y; l.hi[2] = b.z; l.count = __float_as_uint(b.w);
    r.lo[0] = cc.x; r.lo[1] = cc.y; r.lo[2] = cc.z; r.id = __float_as_uint(cc.w);
    r.hi[0] = d.x; r.hi[1] = d.y; r.hi[2] = d.z; r.count = __float_as_uint(d.w);
}

// The collapse only records which primitive lands in which packed slot; k_pack_tris then gathers the vertices of
// all slots at once (one thread per slot: the index -> vertex gather is a chain of dependent DRAM reads that must not
// sit on the collapse's per-node critical path).
struct LeafSinkTriangles {
    PackedTri *tris;
    __device__ __forceinline__ void emit(uint32_t dst, uint32_t prim) const { tris[dst].prim = prim; }
};
struct LeafSinkInstances {
    const uint32_t *active; uint32_t *prim_ids;
    __device__ __forceinline__ void emit(uint32_t dst, uint32_t prim) const { prim_ids[dst] = active[prim]; }
};

// Quantise [lo,hi] of one child against the node frame, conservatively (directed rounding).
__device__ __forceinline__ void quantise_axis(float clo, float chi, float org, float inv_scale, qplane_t &qlo, qplane_t &qhi) {
    float l = floorf(__fmul_rd(__fsub_rd(clo, org), inv_scale));
    float u = ceilf(__fmul_ru(__fsub_ru(chi, org), inv_scale));
    l = fminf(fmaxf(l, 0.f), (float)kQMax);
    u = fminf(fmaxf(u, 0.f), (float)kQMax);
    qlo = (qplane_t)l; qhi = (qplane_t)u;
}

// One 8-lane group per wide node, lane j holding child j in registers: the split search, the node frame, the greedy
// octant slot assignment and the quantisation are 3-step shuffle reductions inside the group instead of serial loops
// over local-memory arrays; the finished node is assembled in shared memory and leaves as one coalesced 128-byte
// store.  The kernel is launched cooperatively (all CTAs co-resident) and walks the wide tree level by level: the
// nodes of a level are a contiguous id range handed out CTA-strided, children and leaf slots are allocated with ONE
// 64-bit atomic per CTA and step (same-address atomics serialise at L2: one per node was the bottleneck), and a grid
// barrier whose last arriver snapshots node_count separates the levels — no polling.
constexpr int kCollapseThreads = 256;
constexpr int kCollapseGroups = kCollapseThreads / 8;

#ifndef LCB_COLLAPSE_MIN_BLOCKS
#define LCB_COLLAPSE_MIN_BLOCKS 5
#endif
template <class Sink>
__global__ void __launch_bounds__(kCollapseThreads, LCB_COLLAPSE_MIN_BLOCKS) k_collapse(const BinNode *__restrict__ bin, const uint32_t *__restrict__ prim_sorted, uint32_t n, BuildHeader *h,
                                                               unsigned long long *queue, WideNode *nodes, uint32_t capacity, Sink sink) {
    __shared__ WideNode s_node[kCollapseGroups];
    __shared__ uint32_t s_int[kCollapseGroups], s_prm[kCollapseGroups];
    __shared__ unsigned long long s_base;
    const uint32_t lane = threadIdx.x & 31, sub = lane & 7u, grp = threadIdx.x >> 3;
    const uint32_t gmask = 0xffu << (lane & 24u);
    const uint32_t below = gmask & ((1u << lane) - 1u);  // lanes of this group below me
    WideNode &out = s_node[grp];
#define GSHFL(V, SRC) __shfl_sync(gmask, V, SRC, 8)
#define GXOR(V, M) __shfl_xor_sync(gmask, V, M, 8)
    uint32_t level_begin = 0, level_end = 1;
    for (uint32_t depth = 0; level_begin < level_end; depth++) {
      const uint32_t level_n = level_end - level_begin;
      for (uint32_t base = blockIdx.x * kCollapseGroups; base < level_n; base += gridDim.x * kCollapseGroups) {
        const bool active = base + grp < level_n;
        const uint32_t t = level_begin + base + grp;
        // ================= phase A: choose the (up to) 8 children and their slots =================
        Child c;  // after phase A: the child of slot `sub`
        for (int k = 0; k < 3; k++) { c.lo[k] = FLT_MAX; c.hi[k] = -FLT_MAX; }
        c.id = 0; c.count = 0;
        bool occupied = false;
        float nlo[3] = {0.f, 0.f, 0.f}; uint32_t ex[3] = {1u, 1u, 1u}; float inv_scale[3] = {0.f, 0.f, 0.f};
        if (active) {
            const uint32_t bnode = (uint32_t)__ldcg(queue + t);  // written by another CTA in the previous level: read at L2
            Child mine = c;
            uint32_t nc;
            if (bnode & kLeafBit) {  // single-primitive tree
                nc = 1;
                if (sub == 0) {
                    for (int k = 0; k < 3; k++) { mine.lo[k] = h->root_lo[k]; mine.hi[k] = h->root_hi[k]; }
                    mine.id = bnode; mine.count = 1;
                }
            } else {
                Child self, l, r; self.id = bnode;
                split_child(bin, self, l, r);
                nc = 2;
                if (sub == 0) mine = l; else if (sub == 1) mine = r;
            }
            // phase 0: open the largest subtree that cannot be a leaf; phase 1: use spare slots to split multi-primitive
            // leaves (tighter boxes at no traversal cost: all 8 slots are tested anyway)
            for (int phase = 0; phase < 2; phase++) {
                const uint32_t limit = phase == 0 ? (uint32_t)kLeafMax : 1u;
                while (nc < 8) {
                    // argmax of the half-area over the group in ONE redux: non-negative floats order like their bit patterns; the low
                    // three mantissa bits carry 7 - lane (ties and near-ties go to the lowest lane)
                    const float a = (sub < nc && mine.count > limit) ? half_area(mine) : -1.f;
                    const uint32_t akey = a >= 0.f ? (((__float_as_uint(a) & ~7u) + 8u) | (7u - sub)) : 0u;
                    const uint32_t abest = __reduce_max_sync(gmask, akey);
                    if (abest == 0u) break;
                    const uint32_t who = 7u - (abest & 7u);
                    Child p, l, r;
                    p.id = GSHFL(mine.id, who);
                    split_child(bin, p, l, r);
                    if (sub == who) mine = l; else if (sub == nc) mine = r;
                    nc++;
                }
            }
            const bool valid = sub < nc;
            // ---- node frame ----
            float nhi[3];
#pragma unroll
            for (int k = 0; k < 3; k++) {
                // group min / max through the order-preserving integer image of the floats (one redux each); invalid lanes hold
                // (+max, -max), NaN is treated the same way (fminf / fmaxf would skip it too)
                const float vlo = mine.lo[k] == mine.lo[k] ? mine.lo[k] : FLT_MAX, vhi = mine.hi[k] == mine.hi[k] ? mine.hi[k] : -FLT_MAX;
                nlo[k] = ordered_to_float(__reduce_min_sync(gmask, float_to_ordered(vlo)));
                nhi[k] = ordered_to_float(__reduce_max_sync(gmask, float_to_ordered(vhi)));
                const float sc = __fdiv_ru(__fsub_ru(nhi[k], nlo[k]), (float)kQMax);
                const uint32_t bits = __float_as_uint(sc);
                uint32_t e = (bits >> 23) + ((bits & 0x7fffffu) ? 1u : 0u);
                e = max(e, 1u); e = min(e, 253u);
                ex[k] = e;
                inv_scale[k] = __uint_as_float((254u - e) << 23);
            }
            // ---- octant slot assignment: slot s is visited first by rays whose direction signs are s (bit k set =
            // negative along axis k); greedy global minimum of dot(child centre - node centre, sign_s), ties -> lowest
            // child, then lowest slot ----
            const float cx = 0.5f * (mine.lo[0] + mine.hi[0]) - 0.5f * (nlo[0] + nhi[0]);
            const float cy = 0.5f * (mine.lo[1] + mine.hi[1]) - 0.5f * (nlo[1] + nhi[1]);
            const float cz = 0.5f * (mine.lo[2] + mine.hi[2]) - 0.5f * (nlo[2] + nhi[2]);
            uint32_t slot_used = 0, my_slot = 8;
            for (uint32_t it = 0; it < nc; it++) {
                float best = FLT_MAX; uint32_t bs = 8;
                if (valid && my_slot == 8) {
#pragma unroll
                    for (uint32_t sl = 0; sl < 8; sl++) {
                        if (slot_used >> sl & 1u) continue;
                        const float cost = ((sl & 1) ? -cx : cx) + ((sl & 2) ? -cy : cy) + ((sl & 4) ? -cz : cz);
                        if (cost < best) { best = cost; bs = sl; }
                    }
                }
                // group argmin in ONE redux: order-preserving image of the cost with its low six bits replaced by (child, slot)
                const uint32_t cbits = __float_as_uint(best);
                const uint32_t cord = cbits ^ ((cbits >> 31) ? 0xffffffffu : 0x80000000u);
                const uint32_t ckey = (valid && my_slot == 8 && bs != 8) ? ((cord & ~63u) | (sub << 3) | bs) : 0xffffffffu;
                const uint32_t cbest = __reduce_min_sync(gmask, ckey);
                uint32_t who = cbest == 0xffffffffu ? 8u : ((cbest >> 3) & 7u);
                if (who != 8u) bs = cbest & 7u;
                if (who == 8u) {  // only NaN costs left: lowest unassigned child takes the lowest free slot
                    who = __ffs(__ballot_sync(gmask, valid && my_slot == 8) >> (lane & 24u)) - 1;
                    bs = __ffs(~slot_used & 0xffu) - 1;
                }
                if (sub == who) my_slot = bs;
                slot_used |= 1u << bs;
            }
            // ---- bring the children into slot order: lane s now holds the child of slot s ----
            uint32_t src = 8;
#pragma unroll
            for (uint32_t j = 0; j < 8; j++) { const uint32_t sj = GSHFL(my_slot, j); if (sj == sub) src = j; }
            const uint32_t from = src & 7u;
#pragma unroll
            for (int k = 0; k < 3; k++) { c.lo[k] = GSHFL(mine.lo[k], from); c.hi[k] = GSHFL(mine.hi[k], from); }
            c.id = GSHFL(mine.id, from); c.count = GSHFL(mine.count, from);
            occupied = src != 8;
        }
        const bool is_int = occupied && c.count > (uint32_t)kLeafMax;
        const bool is_leaf = occupied && !is_int;
        const uint32_t int_ballot = __ballot_sync(gmask, is_int);
        const uint32_t n_internal = __popc(int_ballot), int_rank = __popc(int_ballot & below);
        uint32_t leaf_count = is_leaf ? c.count : 0u, prim_off = leaf_count;
#pragma unroll
        for (int m = 1; m < 8; m <<= 1) { const uint32_t o = __shfl_up_sync(gmask, prim_off, m, 8); if (sub >= (uint32_t)m) prim_off += o; }
        const uint32_t n_prims = GSHFL(prim_off, 7);
        prim_off -= leaf_count;  // exclusive
        // ================= allocation: one 64-bit atomic per CTA and step =================
        if (sub == 0) { s_int[grp] = n_internal; s_prm[grp] = n_prims; }
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t ti = 0, tp = 0;
            for (int g = 0; g < kCollapseGroups; g++) { const uint32_t a = s_int[g], b = s_prm[g]; s_int[g] = ti; s_prm[g] = tp; ti += a; tp += b; }
            s_base = (ti | tp) ? atomicAdd(reinterpret_cast<unsigned long long *>(&h->node_count), (unsigned long long)ti | ((unsigned long long)tp << 32)) : 0ull;
        }
        __syncthreads();
        const uint32_t child_base = (uint32_t)s_base + s_int[grp], prim_base = (uint32_t)(s_base >> 32) + s_prm[grp];
        __syncthreads();  // s_int / s_prm / s_base are rewritten by the next step
        if (!active) continue;
        if (child_base + n_internal > capacity || depth + 1 > (uint32_t)kMaxWideDepth) {
            if (sub == 0) atomicExch(&h->error, child_base + n_internal > capacity ? 2u : 1u);
            continue;  // every CTA still has to reach the barrier; the level loop ends on the error flag
        }
        // ================= phase B: assemble the node in shared memory, store it as one 128-byte line =================
        if (sub == 0) {
            for (int k = 0; k < 3; k++) { out.org[k] = nlo[k]; out.e[k] = (uint8_t)ex[k]; }
            out.imask = (uint8_t)(int_ballot >> (lane & 24u));
            out.child_base = child_base; out.prim_base = prim_base;
        }
        if (!occupied) {
            out.meta[sub] = 0;
            for (int k = 0; k < 3; k++) { out.q[k][0][sub] = (qplane_t)kQMax; out.q[k][1][sub] = 0; }
        } else {
            for (int k = 0; k < 3; k++) quantise_axis(c.lo[k], c.hi[k], nlo[k], inv_scale[k], out.q[k][0][sub], out.q[k][1][sub]);
            if (is_int) {
                out.meta[sub] = (uint8_t)(0x20u | (24u + sub));
                __stcg(queue + child_base + int_rank, (unsigned long long)c.id);
            } else {
                out.meta[sub] = (uint8_t)((((1u << c.count) - 1u) << 5) | prim_off);
                // the (at most kLeafMax) leaves below this child, left to right; valid for any binary tree (LBVH or PLOC)
                uint32_t todo[4], sp = 0, q = 0;
                todo[sp++] = c.id;
                while (sp) {
                    const uint32_t id = todo[--sp];
                    if (id & kLeafBit) sink.emit(prim_base + prim_off + q++, prim_sorted[id & ~kLeafBit]);
                    else {
                        const float4 *pn = reinterpret_cast<const float4 *>(&bin[id]);
                        todo[sp++] = __float_as_uint(pn[2].w); todo[sp++] = __float_as_uint(pn[0].w);
                    }
                }
            }
        }
        __syncwarp(gmask);
        if (sub < (uint32_t)kNodeQuads) reinterpret_cast<uint4 *>(&nodes[t])[sub] = reinterpret_cast<const uint4 *>(&out)[sub];
        __syncwarp(gmask);
      }
      // ---- grid barrier; the last CTA to arrive publishes the end of the next level -------------------------------
      __syncthreads();
      if (threadIdx.x == 0) {
          __threadfence();
          const uint32_t arrived = atomicAdd(&h->bar_count, 1u) + 1u;
          if (arrived == gridDim.x * (depth + 1)) {
              const uint32_t err = *reinterpret_cast<volatile uint32_t *>(&h->error);
              const uint32_t end = *reinterpret_cast<volatile uint32_t *>(&h->node_count);
              h->level_end[depth + 1] = err ? level_end : end;  // an error ends the walk: the next level is empty
              h->max_depth = depth + 1;
              h->emitted = *reinterpret_cast<volatile uint32_t *>(&h->prim_count);
              __threadfence();
              atomicExch(&h->bar_release, depth + 1);
          } else {
              while (*reinterpret_cast<volatile uint32_t *>(&h->bar_release) < depth + 1) __nanosleep(64);
          }
          __threadfence();
      }
      __syncthreads();
      level_begin = level_end;
      level_end = *reinterpret_cast<volatile uint32_t *>(&h->level_end[depth + 1]);
    }
#undef GSHFL
#undef GXOR
}

// One thread per packed slot: gather the slot's triangle (id recorded by the collapse, or kept from the last build when
// refitting) through the index buffer and write the 48 used bytes of the record.
__global__ void __launch_bounds__(256) k_pack_tris(TriangleInput in, PackedTri *tris, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t prim = tris[i].prim;
    float a[3], b[3], c[3];
    load_triangle(in, prim, a, b, c);
    float4 *o = reinterpret_cast<float4 *>(&tris[i]);
    o[0] = make_float4(a[0], a[1], a[2], __uint_as_float(prim));
    o[1] = make_float4(b[0], b[1], b[2], 0.f);
    o[2] = make_float4(c[0], c[1], c[2], 0.f);
}

__global__ void k_seed_queue(const BuildHeader *h, unsigned long long *queue) { queue[0] = (unsigned long long)h->root; }

template <class Sink>
void run_pipeline_after_boxes(cudaStream_t s, uint32_t n, const BuildScratch &sc, WideNode *nodes, uint32_t capacity, const Sink &sink, LaunchCounter &lc, bool ploc) {
    // Morton resolution follows the primitive count: log2(n) bits only enumerate the primitives, the rest resolves non-uniform
    // density — 12 extra bits measured as good as 28 on the bench scenes (profiles/r01r_sort_passes.txt); each pass saved is one
    // sweep over the pairs (24 B per pair).  LC_B200_SORT_PASSES overrides (3..6).
    static const int forced_passes = [] { const char *e = getenv("LC_B200_SORT_PASSES"); int v = e ? atoi(e) : 0; return v >= 3 && v <= 6 ? v : 0; }();
    int passes = forced_passes;
    if (!passes) {
        int lg = 0; while ((1ull << lg) < n) lg++;
        passes = (lg + 12 + 7) / 8;
        passes = passes < 3 ? 3 : (passes > 6 ? 6 : passes);
    }
    k_morton<<<(n + 255) / 256, 256, 0, s>>>(sc.boxes, n, sc.header, sc.keys, sc.vals, 63 - 8 * passes); lc.count++;
    bool in_alt = sort_pairs(s, n, sc.keys, sc.vals, sc.keys_alt, sc.vals_alt, sc.sort_scratch, 0, passes, lc);
    const uint64_t *keys = in_alt ? sc.keys_alt : sc.keys;
    const uint32_t *vals = in_alt ? sc.vals_alt : sc.vals;
    if (ploc && n > 1) {
        static const int radius = [] { int r = 8; if (const char *e = getenv("LC_B200_PLOC_RADIUS")) r = atoi(e); return r < 1 ? 1 : (r > kPlocMaxRadius ? kPlocMaxRadius : r); }();
        const uint32_t tiles = (n + kPlocTile - 1) / kPlocTile;
        float4 *a = sc.ploc_a, *b = sc.ploc_b;
        uint32_t *keep = reinterpret_cast<uint32_t *>(sc.flags);
        k_ploc_init<<<(n + 255) / 256, 256, 0, s>>>(vals, sc.boxes, n, a, sc.ploc_state); lc.count++;
        // Mutual nearest neighbours always exist, so every iteration merges; in practice the cluster count shrinks by 20-45 % per
        // iteration.  The kernels are persistent over tiles and read the live count on the device, so iterations are launched in
        // chunks without the host knowing it; it is read back once per chunk (usually once per build).
        // Adversarial inputs (distances growing monotonically along the order) can leave one mutual pair per iteration: after 72
        // iterations the remaining clusters are paired by position, which halves their number per iteration.
        static int resident = 0;
        if (!resident) {
            int dev = 0, sms = 0, per_sm = 0;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_ploc_step, kPlocTile, 0);
            resident = sms * (per_sm > 0 ? per_sm : 1);
        }
        const uint32_t grid = tiles < (uint32_t)resident ? tiles : (uint32_t)resident;
        uint32_t live = n;
        for (int chunk = 0; live > 1; chunk++) {
            const int iters = chunk == 0 ? 40 : 16;
            const int forced = chunk >= 3 ? 1 : 0;  // after 72 iterations
            for (int it = 0; it < iters; it++) {
                k_ploc_step<<<grid, kPlocTile, 0, s>>>(a, b, keep, sc.ploc_counts, sc.bin, sc.ploc_state, radius, forced);
                k_ploc_scan<<<1, 1024, 0, s>>>(sc.ploc_counts, sc.ploc_state);
                k_ploc_compact<<<grid, kPlocTile, 0, s>>>(b, keep, sc.ploc_counts, a, sc.ploc_state);
                lc.count += 3;
            }
            PlocState st;
            cudaMemcpyAsync(&st, sc.ploc_state, sizeof(st), cudaMemcpyDeviceToHost, s);
            cudaStreamSynchronize(s);
            live = st.n_clusters;
        }
        (void)tiles;
        k_ploc_finish<<<1, 1, 0, s>>>(a, sc.ploc_state, sc.header); lc.count++;
    } else {
        k_hierarchy<<<(n + 127) / 128, 128, 0, s>>>(keys, vals, sc.boxes, n, sc.bin, sc.flags, sc.header); lc.count++;
    }
    // seed the collapse queue with the binary root (device-side, no host round trip)
    k_seed_queue<<<1, 1, 0, s>>>(sc.header, sc.queue); lc.count++;
    // one 8-lane group per wide node of the widest level; cooperative launch: the grid must be co-resident for the barrier
    static int max_blocks = 0;
    if (!max_blocks) {
        int dev = 0, sms = 0, per_sm = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_collapse<Sink>, kCollapseThreads, 0);
        max_blocks = sms * (per_sm > 0 ? per_sm : 1);
    }
    uint32_t blocks = (n / 6 + kCollapseGroups - 1) / kCollapseGroups + 1;
    if (blocks > (uint32_t)max_blocks) blocks = (uint32_t)max_blocks;
    const BinNode *a_bin = sc.bin; const uint32_t *a_vals = vals; uint32_t a_n = n, a_cap = capacity; BuildHeader *a_h = sc.header;
    unsigned long long *a_queue = sc.queue; WideNode *a_nodes = nodes; Sink a_sink = sink;
    void *args[] = {&a_bin, &a_vals, &a_n, &a_h, &a_queue, &a_nodes, &a_cap, &a_sink};
    cudaLaunchCooperativeKernel((const void *)k_collapse<Sink>, dim3(blocks), dim3(kCollapseThreads), args, 0, s); lc.count++;
}

}  // namespace

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

BuildScratch build_scratch_layout(void *base, uint32_t n) {
    BuildScratch sc{};
    uint8_t *p = (uint8_t *)base;
    size_t off = 0;
    auto take = [&](size_t bytes) { void *r = p ? p + off : nullptr; off = align_up(off + bytes, 256); return r; };
    const size_t nn = n ? n : 1;
    sc.header = (BuildHeader *)take(sizeof(BuildHeader));
    sc.boxes = (PrimBox *)take(nn * sizeof(PrimBox));
    sc.keys = (uint64_t *)take(nn * 8);
    sc.keys_alt = (uint64_t *)take(nn * 8);
    sc.vals = (uint32_t *)take(nn * 4);
    sc.vals_alt = (uint32_t *)take(nn * 4);
    sc.sort_scratch = take(sort_scratch_bytes(n, 8));
    sc.bin = (BinNode *)take(nn * sizeof(BinNode));
    sc.flags = (int *)take(nn * 4);
    sc.queue = (unsigned long long *)take(nn * 8);
    sc.ploc_a = (float4 *)take(nn * 32);
    sc.ploc_b = (float4 *)take(nn * 32);
    sc.ploc_counts = (uint32_t *)take((nn / kPlocTile + 2) * 4);
    sc.ploc_state = (PlocState *)take(sizeof(PlocState));
    sc.total_bytes = off;
    return sc;
}

int build_blas(cudaStream_t s, uint32_t n, const TriangleInput &in, const BuildScratch &sc, WideNode *nodes, uint32_t capacity, PackedTri *tris, LaunchCounter &lc, int builder) {
    uint32_t init_blocks = (n + 255) / 256; if (init_blocks > 1024) init_blocks = 1024;
    k_init_header<<<init_blocks, 256, 0, s>>>(sc.header, sc.flags, n); lc.count++;
    k_triangle_boxes<<<(n + 255) / 256, 256, 0, s>>>(in, n, sc.boxes, sc.header); lc.count++;
    bool use_ploc = builder == kBuilderPloc;
    if (builder == kBuilderAuto && n > 1) {
        // PLOC pays off where primitives tile a surface (terrain, C4 / C5: +11 % Mrays/s) and loses to the spatial-median splits of
        // the LBVH where they overlap volumetrically (random soup, C3: -10 %); profiles/r01m_builder_sweep.txt.  The two cases are
        // told apart by how many times the primitives' boxes cover the box of their centroids.
        BuildHeader hdr;
        cudaMemcpyAsync(&hdr, sc.header, sizeof(hdr), cudaMemcpyDeviceToHost, s);
        cudaStreamSynchronize(s);
        float ext[3];
        for (int k = 0; k < 3; k++) ext[k] = ordered_to_float(hdr.bounds_hi[k]) - ordered_to_float(hdr.bounds_lo[k]);
        const float scene = ext[0] * ext[1] + ext[1] * ext[2] + ext[2] * ext[0];
        use_ploc = scene > 0.f && hdr.prim_area_sum < 4.0f * scene;
    }
    LeafSinkTriangles sink{tris};
    run_pipeline_after_boxes(s, n, sc, nodes, capacity, sink, lc, use_ploc);
    k_pack_tris<<<(n + 255) / 256, 256, 0, s>>>(in, tris, n); lc.count++;
    return use_ploc && n > 1 ? kBuilderPloc : kBuilderLbvh;
}

void build_procedural(cudaStream_t s, uint32_t n, const uint8_t *aabbs, const BuildScratch &sc, WideNode *nodes, uint32_t capacity, PackedTri *slots, LaunchCounter &lc) {
    uint32_t init_blocks = (n + 255) / 256; if (init_blocks > 1024) init_blocks = 1024;
    k_init_header<<<init_blocks, 256, 0, s>>>(sc.header, sc.flags, n); lc.count++;
    k_aabb_boxes<<<(n + 255) / 256, 256, 0, s>>>(aabbs, n, sc.boxes, sc.header); lc.count++;
    LeafSinkTriangles sink{slots};
    run_pipeline_after_boxes(s, n, sc, nodes, capacity, sink, lc, false);
    k_pack_aabbs<<<(n + 255) / 256, 256, 0, s>>>(aabbs, slots, n); lc.count++;
}

void build_curves(cudaStream_t s, uint32_t n, const CurveInput &in, const BuildScratch &sc, WideNode *nodes, uint32_t capacity, PackedTri *slots, LaunchCounter &lc) {
    uint32_t init_blocks = (n + 255) / 256; if (init_blocks > 1024) init_blocks = 1024;
    k_init_header<<<init_blocks, 256, 0, s>>>(sc.header, sc.flags, n); lc.count++;
    k_curve_boxes<<<(n + 255) / 256, 256, 0, s>>>(in, n, sc.boxes, sc.header); lc.count++;
    LeafSinkTriangles sink{slots};
    run_pipeline_after_boxes(s, n, sc, nodes, capacity, sink, lc, false);
    k_pack_curves<<<(n + 255) / 256, 256, 0, s>>>(in, slots, n); lc.count++;
    if (in.basis != kCurveLinear) { const uint32_t n_segs = n / in.pieces; k_curve_coefs<<<(n_segs + 255) / 256, 256, 0, s>>>(in, slots, n, n_segs); lc.count++; }
}

void build_tlas(cudaStream_t s, uint32_t n, const uint32_t *active_ids, const InstanceRec *instances, const BuildScratch &sc, WideNode *nodes,
                uint32_t *prim_ids, LaunchCounter &lc) {
    uint32_t init_blocks = (n + 255) / 256; if (init_blocks > 1024) init_blocks = 1024;
    k_init_header<<<init_blocks, 256, 0, s>>>(sc.header, sc.flags, n); lc.count++;
    k_instance_boxes<<<(n + 127) / 128, 128, 0, s>>>(active_ids, n, instances, sc.boxes, sc.header); lc.count++;
    LeafSinkInstances sink{active_ids, prim_ids};
    run_pipeline_after_boxes(s, n, sc, nodes, n, sink, lc, false);  // a handful of instances: the LBVH order is as good as any; the node array holds n
}

// ---- refit (MeshBuild with PreferUpdate on an updatable mesh; GeometryImpl::build_mesh, cpu/accel.rs:251-258) ----
// Topology, slot assignment and primitive order of the wide tree are kept; packed triangles are re-read from the
// (aliased) user vertex buffer and every node's frame and quantised child planes are recomputed bottom-up.  Threads
// start at the nodes without internal children; a node is processed by the last of its internal children to arrive
// (atomic counter), exactly once.
__global__ void __launch_bounds__(256) k_refit_parents(const WideNode *__restrict__ nodes, uint32_t n_nodes, uint32_t *parent, uint32_t *counters) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_nodes) return;
    if (i == 0) parent[0] = 0xffffffffu;
    counters[i] = 0;
    const uint32_t imask = nodes[i].imask, base = nodes[i].child_base;
    const uint32_t n_int = __popc(imask);
    for (uint32_t r = 0; r < n_int; r++) parent[base + r] = i;
}

__global__ void __launch_bounds__(256) k_refit_reset(uint32_t *counters, uint32_t n_nodes) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_nodes) counters[i] = 0;
}

__global__ void __launch_bounds__(128) k_refit(WideNode *nodes, const PackedTri *tris, uint32_t n_nodes, const uint32_t *__restrict__ parent,
                                               float *boxes, uint32_t *counters) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_nodes) return;
    if (nodes[i].imask != 0) return;  // reached later through its children
    while (true) {
        WideNode node = nodes[i];
        float clo[8][3], chi[8][3];
        float nlo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, nhi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
        uint32_t int_rank = 0;
        for (int s = 0; s < 8; s++) {
            const uint32_t meta = node.meta[s];
            if (meta == 0) continue;
            float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
            if (node.imask >> s & 1) {
                const float *b = boxes + 6 * (size_t)(node.child_base + int_rank);
                for (int k = 0; k < 3; k++) { lo[k] = __ldcg(b + k); hi[k] = __ldcg(b + 3 + k); }
                int_rank++;
            } else {
                const uint32_t count = __popc(meta >> 5), first = node.prim_base + (meta & 31u);
                for (uint32_t q = 0; q < count; q++) {
                    const float4 *pt = reinterpret_cast<const float4 *>(&tris[first + q]);  // refreshed by k_pack_tris just before
                    const float4 a = pt[0], b = pt[1], c = pt[2];
                    lo[0] = fminf(lo[0], fmin3(a.x, b.x, c.x)); hi[0] = fmaxf(hi[0], fmax3(a.x, b.x, c.x));
                    lo[1] = fminf(lo[1], fmin3(a.y, b.y, c.y)); hi[1] = fmaxf(hi[1], fmax3(a.y, b.y, c.y));
                    lo[2] = fminf(lo[2], fmin3(a.z, b.z, c.z)); hi[2] = fmaxf(hi[2], fmax3(a.z, b.z, c.z));
                }
            }
            for (int k = 0; k < 3; k++) { clo[s][k] = lo[k]; chi[s][k] = hi[k]; nlo[k] = fminf(nlo[k], lo[k]); nhi[k] = fmaxf(nhi[k], hi[k]); }
        }
        float inv_scale[3];
        for (int k = 0; k < 3; k++) {
            float sc = __fdiv_ru(__fsub_ru(nhi[k], nlo[k]), (float)kQMax);
            uint32_t bits = __float_as_uint(sc);
            uint32_t ex = (bits >> 23) + ((bits & 0x7fffffu) ? 1u : 0u);
            ex = max(ex, 1u); ex = min(ex, 253u);
            node.e[k] = (uint8_t)ex; node.org[k] = nlo[k];
            inv_scale[k] = __uint_as_float((254u - ex) << 23);
        }
        for (int s = 0; s < 8; s++) {
            if (node.meta[s] == 0) continue;
            for (int k = 0; k < 3; k++) quantise_axis(clo[s][k], chi[s][k], nlo[k], inv_scale[k], node.q[k][0][s], node.q[k][1][s]);
        }
        const uint4 *src = reinterpret_cast<const uint4 *>(&node);
        uint4 *dst = reinterpret_cast<uint4 *>(&nodes[i]);
#pragma unroll
        for (int q = 0; q < kNodeQuads; q++) dst[q] = src[q];
        float *b = boxes + 6 * (size_t)i;
        for (int k = 0; k < 3; k++) { __stcg(b + k, nlo[k]); __stcg(b + 3 + k, nhi[k]); }
        const uint32_t p = parent[i];
        if (p == 0xffffffffu) return;
        __threadfence();
        const uint32_t arrived = atomicAdd(&counters[p], 1u) + 1u;
        if (arrived != (uint32_t)__popc(nodes[p].imask)) return;
        __threadfence();
        i = p;
    }
}

void build_refit_arrays(cudaStream_t s, uint32_t n_nodes, const WideNode *nodes, const RefitArrays &ra, LaunchCounter &lc) {
    if (!n_nodes) return;
    k_refit_parents<<<(n_nodes + 255) / 256, 256, 0, s>>>(nodes, n_nodes, ra.parent, ra.counters); lc.count++;
}

void refit_blas(cudaStream_t s, uint32_t n_nodes, uint32_t n_tris, const TriangleInput &in, WideNode *nodes, PackedTri *tris, const RefitArrays &ra,
                BuildHeader *, LaunchCounter &lc) {
    if (!n_nodes || !n_tris) return;
    k_refit_reset<<<(n_nodes + 255) / 256, 256, 0, s>>>(ra.counters, n_nodes); lc.count++;
    k_pack_tris<<<(n_tris + 255) / 256, 256, 0, s>>>(in, tris, n_tris); lc.count++;
    k_refit<<<(n_nodes + 127) / 128, 128, 0, s>>>(nodes, tris, n_nodes, ra.parent, ra.boxes, ra.counters); lc.count++;
}

// ---- instance table scatter ----------------------------------------------------------------
__global__ void __launch_bounds__(128) k_apply_instance_mods(InstanceRec *table, const InstanceModRec *mods, uint32_t n) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const InstanceModRec &m = mods[i];
    InstanceRec &r = table[m.index];
    // `flags` here are already resolved by the host mirror (which applies the reference's
    // ordering rules); the record carries the full new state of the slot.
    for (int k = 0; k < 12; k++) { r.inv[k] = m.inv[k]; r.affine[k] = m.affine[k]; }
    r.nodes = m.nodes; r.tris = m.tris;
    r.visibility = m.visibility; r.user_id = m.user_id; r.flags = m.flags; r.pad = 0;
}

void apply_instance_mods(cudaStream_t s, InstanceRec *table, const InstanceModRec *mods, uint32_t n_mods, LaunchCounter &lc) {
    if (!n_mods) return;
    k_apply_instance_mods<<<(n_mods + 127) / 128, 128, 0, s>>>(table, mods, n_mods); lc.count++;
}

}  // namespace lcb
