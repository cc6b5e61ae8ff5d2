// bvh_build.cu — BVH builder for sm_100a: primitive boxes + centroid bounds, Morton codes (32-bit keys up to 2^25
// primitives), onesweep sort (radix_sort.cu), the binary tree either by fused bottom-up LBVH emission + AABB refit with
// atomic arrival flags (after Apetrei 2014) or by PLOC, and its collapse — a barrier-free work queue — into 8-wide
// 96-byte quantised nodes with packed 64-byte leaf triangles; refit of the wide tree.
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
constexpr unsigned long long kQueueRoot = 1ull << 32;  // collapse work item of the root: (wide depth + 1) << 32 | binary root (written by whoever forms the root)
#ifndef LCB_LEAF_MAX
#define LCB_LEAF_MAX 2
#endif
constexpr int kLeafMax = LCB_LEAF_MAX;  // primitives per leaf child (unary count in 3 bits)

__device__ __forceinline__ float fmin3(float a, float b, float c) { return fminf(a, fminf(b, c)); }
__device__ __forceinline__ float fmax3(float a, float b, float c) { return fmaxf(a, fmaxf(b, c)); }

__global__ void k_init_header(BuildHeader *h, int *flags, uint32_t n_flags, unsigned long long *queue, uint32_t n_queue) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) {
        for (int k = 0; k < 3; k++) { h->bounds_lo[k] = 0x7fffffff; h->bounds_hi[k] = (int)0x80000000; h->root_lo[k] = 0.f; h->root_hi[k] = 0.f; }
        h->root = 0; h->node_count = 1; h->prim_count = 0; h->emitted = 0; h->tickets = 0; h->unused0 = 0; h->max_depth = 0; h->error = 0;
        h->prim_area_sum = 0.f; h->pad2 = 0.f;
    }
    if (i < 24) h->pad_line[i] = 0;  // the host reads the header back whole
    if (i < 23) h->reserved[i] = 0;
    if (i == 0) h->collapse_done = 0;
    for (uint32_t j = i; j < n_flags; j += gridDim.x * blockDim.x) { flags[j] = -1; if (j < n_queue) queue[j] = 0ull; }  // collapse work items: 0 = not published yet (wide nodes <= n_queue)
}

// block-reduce the centroid (shuffles, then one shared-memory round) and fold it into the header's ordered-int
// bounds: 6 atomics per block instead of per warp
// Every box kernel is a grid-stride loop over a bounded grid (kBoxBlocks): a thread folds the centroids of its primitives into a
// running box, the block reduces once, and the header sees 7 atomics per BLOCK OF THE GRID, not per 256 primitives — they all land on
// one cache line of one L2 slice, and at 20 M primitives 550 k of them were most of the kernel (profiles/r02m_build_20M_launches_summary.csv).
constexpr uint32_t kBoxBlocks = 148 * 8;
struct CentroidAcc {
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX}, area = 0.f;
    __device__ __forceinline__ void add(const float c[3], float a) {
#pragma unroll
        for (int k = 0; k < 3; k++) { lo[k] = fminf(lo[k], c[k]); hi[k] = fmaxf(hi[k], c[k]); }
        area += a;
    }
};
__device__ __forceinline__ void reduce_centroid_bounds(const CentroidAcc &acc, BuildHeader *h) {
    __shared__ float s_lo[8][3], s_hi[8][3], s_area[8];
    float lo[3], hi[3], area = acc.area;
#pragma unroll
    for (int k = 0; k < 3; k++) { lo[k] = acc.lo[k]; hi[k] = acc.hi[k]; }
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
    CentroidAcc acc;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float a[3], b[3], c[3], cen[3];
        load_triangle(in, i, a, b, c);
        PrimBox pb;
#pragma unroll
        for (int k = 0; k < 3; k++) {
            pb.lo[k] = fmin3(a[k], b[k], c[k]);
            pb.hi[k] = fmax3(a[k], b[k], c[k]);
            cen[k] = 0.5f * pb.lo[k] + 0.5f * pb.hi[k];
        }
        reinterpret_cast<float4 *>(boxes)[2 * (size_t)i] = make_float4(pb.lo[0], pb.lo[1], pb.lo[2], 0.f);
        reinterpret_cast<float4 *>(boxes)[2 * (size_t)i + 1] = make_float4(pb.hi[0], pb.hi[1], pb.hi[2], 0.f);
        const float dx = pb.hi[0] - pb.lo[0], dy = pb.hi[1] - pb.lo[1], dz = pb.hi[2] - pb.lo[2];
        acc.add(cen, dx * dy + dy * dz + dz * dx);
    }
    reduce_centroid_bounds(acc, h);
}

// Procedural primitives (ProceduralPrimitiveBuild, api_types:633-641): the user's AABBs {min[3], max[3]} (rtx.rs:339-345, 24 B) are
// the primitive boxes (GeometryImpl::build_procedural's bounds_func, cpu/accel.rs:93-104).
__global__ void __launch_bounds__(256) k_aabb_boxes(const uint8_t *__restrict__ aabbs, uint32_t n, PrimBox *boxes, BuildHeader *h) {
    CentroidAcc acc;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float *a = reinterpret_cast<const float *>(aabbs + (size_t)i * 24);
        float lo[3], hi[3], cen[3];
#pragma unroll
        for (int k = 0; k < 3; k++) { lo[k] = fminf(a[k], a[3 + k]); hi[k] = fmaxf(a[k], a[3 + k]); cen[k] = 0.5f * lo[k] + 0.5f * hi[k]; }
        reinterpret_cast<float4 *>(boxes)[2 * (size_t)i] = make_float4(lo[0], lo[1], lo[2], 0.f);
        reinterpret_cast<float4 *>(boxes)[2 * (size_t)i + 1] = make_float4(hi[0], hi[1], hi[2], 0.f);
        const float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
        acc.add(cen, dx * dy + dy * dz + dz * dx);
    }
    reduce_centroid_bounds(acc, h);
}

// leaf records of a procedural BLAS: the primitive's box and id in the PackedTri slot the collapse assigned
__global__ void __launch_bounds__(256) k_pack_aabbs(const uint8_t *__restrict__ aabbs, const uint32_t *__restrict__ slot_prim, PackedTri *tris, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t prim = slot_prim[i];
    const float *a = reinterpret_cast<const float *>(aabbs + (size_t)prim * 24);
    float4 *o = reinterpret_cast<float4 *>(&tris[i]);
    o[0] = make_float4(fminf(a[0], a[3]), fminf(a[1], a[4]), fminf(a[2], a[5]), __uint_as_float(prim));
    o[1] = make_float4(fmaxf(a[0], a[3]), fmaxf(a[1], a[4]), fmaxf(a[2], a[5]), 0.f);
    o[2] = make_float4(0.f, 0.f, 0.f, 0.f);
}

// Curve pieces (CurveBuild): box of the two end spheres, rounded outward.
__global__ void __launch_bounds__(256) k_curve_boxes(CurveInput in, uint32_t n, PrimBox *boxes, BuildHeader *h) {
    CentroidAcc acc;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float4 A, B;
        curve_piece(in.cps, in.cp_stride, in.segs, in.basis, i / in.pieces, i % in.pieces, A, B);
        const float ra = fabsf(A.w), rb = fabsf(B.w);
        const float pa[3] = {A.x, A.y, A.z}, pb[3] = {B.x, B.y, B.z};
        float lo[3], hi[3], cen[3];
        // the canonical cone test decides in fp32: pad by a fraction of the radius and of the piece's extent so that a ray it
        // accepts never misses the box
        const float pad = fmaxf(ra, rb) * (1.0f / 256.0f) + fmaxf(fmaxf(fabsf(pb[0] - pa[0]), fabsf(pb[1] - pa[1])), fabsf(pb[2] - pa[2])) * (1.0f / 4096.0f);
#pragma unroll
        for (int k = 0; k < 3; k++) {
            lo[k] = __fsub_rd(fminf(__fsub_rd(pa[k], ra), __fsub_rd(pb[k], rb)), pad);
            hi[k] = __fadd_ru(fmaxf(__fadd_ru(pa[k], ra), __fadd_ru(pb[k], rb)), pad);
            cen[k] = 0.5f * lo[k] + 0.5f * hi[k];
        }
        reinterpret_cast<float4 *>(boxes)[2 * (size_t)i] = make_float4(lo[0], lo[1], lo[2], 0.f);
        reinterpret_cast<float4 *>(boxes)[2 * (size_t)i + 1] = make_float4(hi[0], hi[1], hi[2], 0.f);
        const float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
        acc.add(cen, dx * dy + dy * dz + dz * dx);
    }
    reduce_centroid_bounds(acc, h);
}

// leaf records of a curve BLAS: the piece's two spheres, its segment and its parameter range (CurveSeg)
__global__ void __launch_bounds__(256) k_pack_curves(CurveInput in, const uint32_t *__restrict__ slot_prim, PackedTri *slots, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t j = slot_prim[i], seg = j / in.pieces, k = j % in.pieces;
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
    CentroidAcc acc;
    if (valid) acc.add(cen, 0.f);
    reduce_centroid_bounds(acc, h);
}

// orders this thread's earlier stores before its next atomic at GPU scope; what __threadfence() gives beyond that (sequential
// consistency: MEMBAR.SC) is not needed by the arrival protocols here, whose readers go to L2 (__ldcg)
__device__ __forceinline__ void release_fence() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }

__device__ __forceinline__ uint64_t expand21(uint32_t x) {
    uint64_t v = x & 0x1fffffull;
    v = (v | v << 32) & 0x1f00000000ffffull;
    v = (v | v << 16) & 0x1f0000ff0000ffull;
    v = (v | v << 8) & 0x100f00f00f00f00full;
    v = (v | v << 4) & 0x10c30c30c30c30c3ull;
    v = (v | v << 2) & 0x1249249249249249ull;
    return v;
}

template <typename K>
__global__ void __launch_bounds__(256) k_morton(const PrimBox *__restrict__ boxes, uint32_t n, const BuildHeader *__restrict__ h,
                                                K *__restrict__ keys, uint32_t *__restrict__ vals, int drop_bits) {
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
    keys[i] = (K)((expand21(q[0]) | (expand21(q[1]) << 1) | (expand21(q[2]) << 2)) >> drop_bits);
    vals[i] = i;
}

// ---- fused hierarchy + refit ---------------------------------------------------------------
// Internal node i separates sorted leaves i and i+1.  delta(i) orders the splits: the XOR of
// adjacent keys, with runs of equal keys split by position bits.
template <typename K>
__device__ __forceinline__ uint64_t split_delta(const K *__restrict__ keys, uint32_t i) {
    uint64_t x = keys[i] ^ keys[i + 1];
    return x ? (x | (1ull << 63)) : (uint64_t)(i ^ (i + 1));
}

// Two phases.  (1) Warp-local: the 32 consecutive sorted leaves of a warp are merged with shuffles only — every lane owns the
// cluster that starts at its leaf; per step a cluster whose parent lies to its right merges with the next active cluster when
// that one's parent lies to its left (then both are children of internal node `right`); a cluster whose sibling is outside the
// warp's span leaves for phase 2.  31 of 32 internal nodes are emitted here without atomics, fences or L2 round trips.
// (2) Global: the classic bottom-up climb with one atomic exchange per arrival (Apetrei 2014): the first child to arrive at an
// internal node leaves its box there and stops, the second merges and continues.
template <typename K>
__global__ void __launch_bounds__(128) k_hierarchy(const K *__restrict__ keys, const uint32_t *__restrict__ prim, const PrimBox *__restrict__ boxes,
                                                   uint32_t n, BinNode *bin, int *flags, BuildHeader *h, unsigned long long *queue) {
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
            h->root = cur; queue[0] = kQueueRoot | cur;
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
                h->root = parent; queue[0] = kQueueRoot | parent;
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
    // The climb is a chain of dependent L2 round trips per level, and the thread that arrives second is the critical path of the whole
    // kernel.  It first LOOKS at the flag: when the sibling is already there (it fenced before it published the flag, so its half of the
    // node is visible at L2) the level costs two dependent loads — flag, then the sibling's box together with the keys that decide the
    // next level's parent — instead of store, fence, exchange, load.  Only a thread that finds the flag empty runs the full protocol.
    bool parent_on_right = (left == 0) || (right != n - 1 && split_delta(keys, right) < split_delta(keys, left - 1));
    while (true) {
        const uint32_t count = right - left + 1;
        const uint32_t parent = parent_on_right ? right : left - 1;   // left child of internal node `right` / right child of `left - 1`
        float4 *pn = reinterpret_cast<float4 *>(&bin[parent]);
        float4 *mine = parent_on_right ? pn : pn + 2;
        const float4 *theirs = parent_on_right ? pn + 2 : pn;
        const int my_end = parent_on_right ? (int)left : (int)right;
        mine[0] = make_float4(lo.x, lo.y, lo.z, __uint_as_float(cur));
        mine[1] = make_float4(hi.x, hi.y, hi.z, __uint_as_float(count));
        int other = __ldcg(&flags[parent]);
        if (other == -1) {
            release_fence();
            other = atomicExch(&flags[parent], my_end);
            if (other == -1) return;
        }
        if (parent_on_right) right = (uint32_t)other; else left = (uint32_t)other;
        const float4 slo = __ldcg(theirs), shi = __ldcg(theirs + 1);
        const bool is_root = left == 0 && right == n - 1;
        if (!is_root) parent_on_right = (left == 0) || (right != n - 1 && split_delta(keys, right) < split_delta(keys, left - 1));  // issued before the box is consumed
        lo.x = fminf(lo.x, slo.x); lo.y = fminf(lo.y, slo.y); lo.z = fminf(lo.z, slo.z);
        hi.x = fmaxf(hi.x, shi.x); hi.y = fmaxf(hi.y, shi.y); hi.z = fmaxf(hi.z, shi.z);
        cur = parent;
        if (is_root) {
            h->root = parent; queue[0] = kQueueRoot | parent;
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

__global__ void k_ploc_finish(const float4 *clusters, const PlocState *st, BuildHeader *h, unsigned long long *queue) {
    if (st->n_clusters != 1) { h->error = 3u; return; }
    const float4 lo = clusters[0], hi = clusters[1];
    h->root = __float_as_uint(lo.w); queue[0] = kQueueRoot | __float_as_uint(lo.w);
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
    uint32_t *slot_prim;  // compact (4 bytes per slot, not one word of every 64-byte record: that is a sector read-modify-write per slot)
    __device__ __forceinline__ void emit(uint32_t dst, uint32_t prim) const { slot_prim[dst] = prim; }
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
// over local-memory arrays; the finished node is assembled in shared memory and leaves as coalesced 16-byte stores.
//
// Barrier-free work queue (round 1 walked the wide tree level by level with a grid barrier per level).  queue[t] is the
// work item of wide node t: (wide depth + 1) << 32 | binary subtree root — zero until the parent has written it.  Warps are
// autonomous, four groups each: a group takes node ids as TICKETS and polls its entry; a parent allocates its internal
// children as a contiguous id range and writes their entries (two atomics per warp and step: nodes + leaf slots in one
// 64-bit word, tickets).  Tickets are only held by running groups and a node's parent always has a smaller ticket, so the
// queue drains without co-residency requirements.  The walk is over when every primitive has a leaf slot: the allocation
// that places the last one publishes the final node count in `collapse_done` (its own cache line: thousands of idle warps
// poll it, and must not queue up in front of the allocation atomics).
//
// What a wide node costs is a chain of dependent loads from a 64 MB array in DRAM, so the chain is kept short:
//  * every lane PRELOADS the binary node of the child it holds: the opened child's two children arrive by shuffles and only
//    the two lanes that received them issue new loads — the chain is the depth of the opening tree (3-4), not six opens;
//  * SMALL SUBTREES IN ONE ROUND TRIP: in the LBVH numbering an internal node's id is the position of its split, so the
//    subtree over sorted leaves [l, r] owns exactly the ids l .. r-1.  A root with <= 16 leaves (every bottom node, 4 of 5
//    nodes) fetches all of them, and its primitive ids, with independent loads into shared memory / registers and opens
//    from there; PLOC numbers nodes in merge order and takes the general path;
//  * a parent prefetches into L2 what each internal child will touch first (its subtree block, or its children's nodes).
#ifndef LCB_COLLAPSE_THREADS
#define LCB_COLLAPSE_THREADS 128
#endif
constexpr int kCollapseThreads = LCB_COLLAPSE_THREADS;
constexpr int kCollapseGroups = kCollapseThreads / 8;
constexpr uint32_t kSmallSubtree = 16;  // leaves; 15 binary nodes = 960 bytes of shared memory per group
static_assert(kLeafMax <= 2, "k_collapse reads the primitives of a leaf child from the child's preloaded binary node");

// reductions over the 8-lane group of a lane: xor butterflies on the FULL warp mask (the four groups of a warp run in lockstep — a
// redux.sync or shfl.sync on a partial, lane-dependent mask compiles to a MATCH.ANY loop over the distinct masks, ~50 instructions each)
#define GSHFL(V, SRC) __shfl_sync(0xffffffffu, V, SRC, 8)
#define GXOR(V, M) __shfl_xor_sync(0xffffffffu, V, M, 8)
__device__ __forceinline__ uint32_t group_max(uint32_t v) { v = max(v, GXOR(v, 1)); v = max(v, GXOR(v, 2)); return max(v, GXOR(v, 4)); }
__device__ __forceinline__ uint32_t group_min(uint32_t v) { v = min(v, GXOR(v, 1)); v = min(v, GXOR(v, 2)); return min(v, GXOR(v, 4)); }
__device__ __forceinline__ int group_min(int v) { v = min(v, GXOR(v, 1)); v = min(v, GXOR(v, 2)); return min(v, GXOR(v, 4)); }
__device__ __forceinline__ int group_max(int v) { v = max(v, GXOR(v, 1)); v = max(v, GXOR(v, 2)); return max(v, GXOR(v, 4)); }
__device__ __forceinline__ uint32_t group_or(uint32_t v) { v |= GXOR(v, 1); v |= GXOR(v, 2); return v | GXOR(v, 4); }
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

#ifndef LCB_COLLAPSE_MIN_BLOCKS
#define LCB_COLLAPSE_MIN_BLOCKS (1024 / LCB_COLLAPSE_THREADS)
#endif
template <class Sink, bool kContiguous>
__global__ void __launch_bounds__(kCollapseThreads, LCB_COLLAPSE_MIN_BLOCKS) k_collapse(const BinNode *__restrict__ bin, const uint32_t *__restrict__ prim_sorted, uint32_t n, BuildHeader *h,
                                                               unsigned long long *queue, WideNode *nodes, uint32_t capacity, Sink sink) {
    __shared__ WideNode s_node[kCollapseGroups];
    __shared__ float4 s_bin[kContiguous ? kCollapseGroups : 1][(kSmallSubtree - 1) * 4];
    const uint32_t lane = threadIdx.x & 31, sub = lane & 7u, grp = threadIdx.x >> 3, wgrp = lane >> 3;
    const uint32_t gmask = 0xffu << (lane & 24u);
    WideNode &out = s_node[grp];
    float4 *blk = s_bin[kContiguous ? grp : 0];
    // warps are autonomous (no CTA barrier anywhere): four groups, four tickets
    uint32_t t = 0;
    if (lane == 0) t = atomicAdd(&h->tickets, 4u);
    t = __shfl_sync(0xffffffffu, t, 0) + wgrp;  // my ticket = the wide node this group emits next
    bool exhausted = false;                     // no node will ever appear at my ticket
    uint32_t my_depth = 0, my_emitted = 0, idle_streak = 0;
    for (;;) {
        // ================= take the work item, if its parent has published it =================
        unsigned long long item = 0ull;
        if (!exhausted && sub == 0 && t < n && t < capacity) item = __ldcg(queue + t);  // min(n, capacity) entries are cleared: no wide node has an id beyond either
        item = GSHFL(item, 0);
        const bool active = item != 0ull;
        const uint32_t bnode = (uint32_t)item, depth = (uint32_t)(item >> 32) - 1u;
        if (!__any_sync(0xffffffffu, active)) {
            // nothing published for this warp: is the walk over?  collapse_done = final node count + 1 (1 after an error: everybody leaves)
            uint32_t done = 0u;
            if (lane == 0) done = *reinterpret_cast<volatile uint32_t *>(&h->collapse_done);
            done = __shfl_sync(0xffffffffu, done, 0);
            if (done != 0u && t >= done - 1u) exhausted = true;
            if (__all_sync(0xffffffffu, exhausted)) break;
            __nanosleep(128u << min(idle_streak, 3u));
            idle_streak++;
            continue;
        }
        idle_streak = 0;
        // ================= phase A: choose the (up to) 8 children and their slots =================
        Child mine;  // the child this lane holds (children are NOT brought into slot order: the lane writes to its slot's place)
        for (int k = 0; k < 3; k++) { mine.lo[k] = FLT_MAX; mine.hi[k] = -FLT_MAX; }
        mine.id = 0; mine.count = 0;
        uint32_t nc = 0;
        float4 pa = make_float4(0.f, 0.f, 0.f, 0.f), pb = pa, pc = pa, pd = pa;  // preloaded binary node of `mine`
        bool small = false;              // the whole subtree sits in `blk` (uniform over the group)
        uint32_t first = 0;              // small: first sorted leaf = first binary node id of the subtree
        uint32_t prims_lo = 0, prims_hi = 0;  // small: primitive ids of sorted leaves first + sub, first + 8 + sub
        if (active) {
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
                const uint32_t total = l.count + r.count;
                if (kContiguous && total <= kSmallSubtree) {
                    small = true;
                    first = bnode + 1u - l.count;
                    const float4 *srcq = reinterpret_cast<const float4 *>(&bin[first]);
                    const uint32_t quads = (total - 1u) * 4u;
#pragma unroll
                    for (uint32_t q = 0; q < (kSmallSubtree - 1) * 4 / 8 + 1; q++) { const uint32_t qi = q * 8u + sub; if (qi < quads) blk[qi] = srcq[qi]; }
                    if (sub < total) prims_lo = prim_sorted[first + sub];
                    if (8u + sub < total) prims_hi = prim_sorted[first + 8u + sub];
                }
            }
        }
        __syncwarp();  // blk is read by other lanes of the group
        auto fetch = [&](uint32_t id) {
            const float4 *pn = (kContiguous && small) ? blk + (size_t)(id - first) * 4 : reinterpret_cast<const float4 *>(&bin[id]);
            pa = pn[0]; pb = pn[1]; pc = pn[2]; pd = pn[3];
        };
        if (active && sub < nc && !(mine.id & kLeafBit)) fetch(mine.id);
        // Open the largest subtree that cannot be a leaf until there are 8 children; then use spare slots to split two-primitive
        // leaves (tighter boxes at no traversal cost: all 8 slots are tested anyway).  One loop: bit 31 of the key ranks the first
        // kind above the second.  argmax of the half-area: non-negative floats order like their bit patterns; the low three
        // mantissa bits carry 7 - lane (ties and near-ties go to the lowest lane).
        for (;;) {
            uint32_t akey = 0u;
            if (sub < nc && mine.count > 1u) {
                const float a = half_area(mine);
                if (a >= 0.f) akey = (((__float_as_uint(a) & 0x7ffffff8u) + 8u) | (7u - sub)) | (mine.count > (uint32_t)kLeafMax ? 0x80000000u : 0u);  // a NaN box is never opened
            }
            const uint32_t abest = group_max(akey);
            const bool open = nc < 8u && abest != 0u;  // uniform over the group (nc = 0 in a group without work)
            if (!__any_sync(0xffffffffu, open)) break;
            const uint32_t who = 7u - (abest & 7u);
            // the opened child's left child stays in lane `who` (own registers), its right child goes to lane nc
            const float4 qc = make_float4(GSHFL(pc.x, who), GSHFL(pc.y, who), GSHFL(pc.z, who), GSHFL(pc.w, who));
            const float4 qd = make_float4(GSHFL(pd.x, who), GSHFL(pd.y, who), GSHFL(pd.z, who), GSHFL(pd.w, who));
            if (open) {
                const bool take_l = sub == who, take_r = sub == nc;
                if (take_l) { mine.lo[0] = pa.x; mine.lo[1] = pa.y; mine.lo[2] = pa.z; mine.id = __float_as_uint(pa.w); mine.hi[0] = pb.x; mine.hi[1] = pb.y; mine.hi[2] = pb.z; mine.count = __float_as_uint(pb.w); }
                if (take_r) { mine.lo[0] = qc.x; mine.lo[1] = qc.y; mine.lo[2] = qc.z; mine.id = __float_as_uint(qc.w); mine.hi[0] = qd.x; mine.hi[1] = qd.y; mine.hi[2] = qd.z; mine.count = __float_as_uint(qd.w); }
                if ((take_l || take_r) && !(mine.id & kLeafBit)) fetch(mine.id);
                nc++;
            }
        }
        const bool valid = sub < nc;
        const bool is_int = valid && mine.count > (uint32_t)kLeafMax, is_leaf = valid && !is_int;
        // the primitives of a leaf child (a two-primitive child is a binary node over two leaves: preloaded)
        uint32_t pos_a = 0, pos_b = 0, prim_a = 0, prim_b = 0;
        if (is_leaf) {
            if (mine.id & kLeafBit) pos_a = mine.id & ~kLeafBit;
            else { pos_a = __float_as_uint(pa.w) & ~kLeafBit; pos_b = __float_as_uint(pc.w) & ~kLeafBit; }
        }
        if (kContiguous && __any_sync(0xffffffffu, small)) {  // small subtrees: the ids are in the group's registers
            const uint32_t ia = (pos_a - first) & 15u, ib = (pos_b - first) & 15u;
            const uint32_t a_lo = GSHFL(prims_lo, ia & 7u), a_hi = GSHFL(prims_hi, ia & 7u), b_lo = GSHFL(prims_lo, ib & 7u), b_hi = GSHFL(prims_hi, ib & 7u);
            if (small) { prim_a = (ia & 8u) ? a_hi : a_lo; prim_b = (ib & 8u) ? b_hi : b_lo; }
        }
        if (is_leaf && !small) { prim_a = prim_sorted[pos_a]; if (mine.count > 1u) prim_b = prim_sorted[pos_b]; }
        // what an internal child's group will touch first, on its way into L2 while this node is being finished
        if (is_int) {
            if (kContiguous && mine.count <= kSmallSubtree) {
                const uint32_t cfirst = mine.id + 1u - __float_as_uint(pb.w);
                const char *p0 = reinterpret_cast<const char *>(&bin[cfirst]), *p1 = reinterpret_cast<const char *>(&bin[cfirst + mine.count - 1u]);
                for (const char *p = p0; p < p1; p += 128) prefetch_l2(p);
                prefetch_l2(p1 - 1);
                prefetch_l2(prim_sorted + cfirst); prefetch_l2(prim_sorted + cfirst + mine.count - 1u);
            } else {
                if (!(__float_as_uint(pa.w) & kLeafBit)) prefetch_l2(&bin[__float_as_uint(pa.w)]);
                if (!(__float_as_uint(pc.w) & kLeafBit)) prefetch_l2(&bin[__float_as_uint(pc.w)]);
            }
        }
        // ---- node frame ----
        float nlo[3], nhi[3], inv_scale[3]; uint32_t ex[3];
#pragma unroll
        for (int k = 0; k < 3; k++) {
            // group min / max through the order-preserving integer image of the floats; invalid lanes hold (+max, -max), NaN is
            // treated the same way (fminf / fmaxf would skip it too)
            const float vlo = mine.lo[k] == mine.lo[k] ? mine.lo[k] : FLT_MAX, vhi = mine.hi[k] == mine.hi[k] ? mine.hi[k] : -FLT_MAX;
            nlo[k] = ordered_to_float(group_min(float_to_ordered(vlo)));
            nhi[k] = ordered_to_float(group_max(float_to_ordered(vhi)));
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
        float cost[8];
#pragma unroll
        for (uint32_t sl = 0; sl < 8; sl++) cost[sl] = ((sl & 1) ? -cx : cx) + ((sl & 2) ? -cy : cy) + ((sl & 4) ? -cz : cz);
        uint32_t slot_used = 0, my_slot = 8;
        for (uint32_t it = 0; it < 8u; it++) {
            float best = FLT_MAX; uint32_t bs = 8;
            const bool waiting = valid && my_slot == 8;
            if (waiting) {
#pragma unroll
                for (uint32_t sl = 0; sl < 8; sl++) if (!(slot_used >> sl & 1u) && cost[sl] < best) { best = cost[sl]; bs = sl; }
            }
            const uint32_t unassigned = (__ballot_sync(0xffffffffu, waiting) >> (lane & 24u)) & 0xffu;  // uniform over the group
            if (!__any_sync(0xffffffffu, unassigned != 0u)) break;
            // Shortcut: when the waiting children all prefer different free slots (and none is left with NaN costs only), the greedy
            // hands each of them exactly that slot whatever the order — taking one never removes another's favourite — so they are all
            // assigned at once.  Children spread over the octants of their parent: this ends most nodes after one or two rounds.
            const uint32_t wanted = group_or(waiting && bs != 8 ? 1u << bs : 0u);
            const bool all_distinct = (uint32_t)__popc(wanted) == (uint32_t)__popc(unassigned);
            // group argmin: order-preserving image of the cost with its low six bits replaced by (child, slot)
            const uint32_t cbits = __float_as_uint(best);
            const uint32_t cord = cbits ^ ((cbits >> 31) ? 0xffffffffu : 0x80000000u);
            const uint32_t ckey = (waiting && bs != 8) ? ((cord & ~63u) | (sub << 3) | bs) : 0xffffffffu;
            const uint32_t cbest = group_min(ckey);
            if (all_distinct) {
                if (waiting) my_slot = bs;
                slot_used |= wanted;
            } else if (unassigned != 0u) {
                uint32_t who = (cbest >> 3) & 7u, ws = cbest & 7u;
                if (cbest == 0xffffffffu) {  // only NaN costs left: lowest unassigned child takes the lowest free slot
                    who = __ffs(unassigned) - 1;
                    ws = __ffs(~slot_used & 0xffu) - 1;
                }
                if (sub == who) my_slot = ws;
                slot_used |= 1u << ws;
            }
        }
        // ---- per-slot summaries: which slots hold internal children, how many primitives each leaf slot holds (2 bits per slot) ----
        const uint32_t imask = group_or(is_int ? 1u << my_slot : 0u);
        const uint32_t counts = group_or(is_leaf ? mine.count << (2u * my_slot) : 0u);
        const uint32_t lower = (1u << (2u * (my_slot & 7u))) - 1u;
        const uint32_t n_internal = __popc(imask), int_rank = __popc(imask & ((1u << (my_slot & 7u)) - 1u));
        const uint32_t n_prims = __popc(counts & 0x5555u) + 2u * __popc(counts & 0xaaaau);
        const uint32_t prim_off = __popc(counts & lower & 0x5555u) + 2u * __popc(counts & lower & 0xaaaau);
        // ================= allocation + next tickets: two atomics per warp and step =================
        const uint32_t mine_cnt = active ? (1u | (n_internal << 8) | (n_prims << 18)) : 0u;  // uniform over the group
        const uint32_t c0 = __shfl_sync(0xffffffffu, mine_cnt, 0), c1 = __shfl_sync(0xffffffffu, mine_cnt, 8), c2 = __shfl_sync(0xffffffffu, mine_cnt, 16), c3 = __shfl_sync(0xffffffffu, mine_cnt, 24);
        const uint32_t total = c0 + c1 + c2 + c3, excl = (wgrp > 0 ? c0 : 0u) + (wgrp > 1 ? c1 : 0u) + (wgrp > 2 ? c2 : 0u);
        unsigned long long base = 0ull; uint32_t ticket = 0u;
        if (lane == 0) {
            const uint32_t ti = (total >> 8) & 0x3ffu, tp = total >> 18;
            base = atomicAdd(reinterpret_cast<unsigned long long *>(&h->node_count), (unsigned long long)ti | ((unsigned long long)tp << 32));
            ticket = atomicAdd(&h->tickets, total & 0xffu);
            // the allocation that places the last primitive ends the walk: nothing allocated before it is unprocessed (an unprocessed
            // node still holds unplaced primitives), so the node count it sees is final
            if ((uint32_t)(base >> 32) + tp == n) *reinterpret_cast<volatile uint32_t *>(&h->collapse_done) = (uint32_t)base + ti + 1u;
        }
        base = __shfl_sync(0xffffffffu, base, 0); ticket = __shfl_sync(0xffffffffu, ticket, 0);
        const uint32_t child_base = (uint32_t)base + ((excl >> 8) & 0x3ffu), prim_base = (uint32_t)(base >> 32) + (excl >> 18);
        const uint32_t next_ticket = ticket + (excl & 0xffu);
        if (active) {
            if (child_base + n_internal > capacity || depth + 1 > (uint32_t)kMaxWideDepth) {
                if (sub == 0) {  // every group leaves at its next idle step
                    atomicExch(&h->error, child_base + n_internal > capacity ? 2u : 1u);
                    *reinterpret_cast<volatile uint32_t *>(&h->collapse_done) = 1u;
                }
            } else {
                // ================= phase B: assemble the node in shared memory, store it in 16-byte pieces =================
                out.meta[sub] = 0;  // every slot empty; the lanes that hold a child then fill theirs
                for (int k = 0; k < 3; k++) { out.q[k][0][sub] = (qplane_t)kQMax; out.q[k][1][sub] = 0; }
                if (sub == 0) {
                    for (int k = 0; k < 3; k++) { out.org[k] = nlo[k]; out.e[k] = (uint8_t)ex[k]; }
                    out.imask = (uint8_t)imask;
                    out.child_base = child_base; out.prim_base = prim_base;
                }
                __syncwarp(gmask);
                if (valid) {
                    for (int k = 0; k < 3; k++) quantise_axis(mine.lo[k], mine.hi[k], nlo[k], inv_scale[k], out.q[k][0][my_slot], out.q[k][1][my_slot]);
                    if (is_int) {
                        out.meta[my_slot] = (uint8_t)(0x20u | (24u + my_slot));
                        __stcg(queue + child_base + int_rank, ((unsigned long long)(depth + 2u) << 32) | (unsigned long long)mine.id);
                    } else {
                        out.meta[my_slot] = (uint8_t)((((1u << mine.count) - 1u) << 5) | prim_off);
                        sink.emit(prim_base + prim_off, prim_a);
                        if (mine.count > 1u) sink.emit(prim_base + prim_off + 1u, prim_b);
                        my_emitted += mine.count;
                    }
                }
                __syncwarp(gmask);
                if (sub < (uint32_t)kNodeQuads) reinterpret_cast<uint4 *>(&nodes[t])[sub] = reinterpret_cast<const uint4 *>(&out)[sub];
                __syncwarp(gmask);
                my_depth = max(my_depth, depth + 1u);
            }
            t = next_ticket;
        }
    }
    // ---- per-warp totals: deepest node, primitives emitted ----
    for (int off = 16; off > 0; off >>= 1) { my_depth = max(my_depth, __shfl_xor_sync(0xffffffffu, my_depth, off)); my_emitted += __shfl_xor_sync(0xffffffffu, my_emitted, off); }
    if (lane == 0) { if (my_depth) atomicMax(&h->max_depth, my_depth); if (my_emitted) atomicAdd(&h->emitted, my_emitted); }
}

// One thread per packed slot: gather the slot's triangle (id recorded by the collapse, or kept from the last build when
// refitting) through the index buffer and write the record.
__global__ void __launch_bounds__(256) k_pack_tris(TriangleInput in, const uint32_t *__restrict__ slot_prim, PackedTri *tris, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t prim = slot_prim ? slot_prim[i] : tris[i].prim;  // build: the collapse's slot -> primitive list; refit: kept in the record
    float a[3], b[3], c[3];
    load_triangle(in, prim, a, b, c);
    float4 *o = reinterpret_cast<float4 *>(&tris[i]);
    o[0] = make_float4(a[0], a[1], a[2], __uint_as_float(prim));
    o[1] = make_float4(b[0], b[1], b[2], 0.f);
    o[2] = make_float4(c[0], c[1], c[2], 0.f);
    o[3] = make_float4(0.f, 0.f, 0.f, 0.f);  // the spare quad too: two whole sectors per record instead of a read-modify-write of the second
}


static inline uint32_t box_blocks(uint32_t n) { const uint32_t b = (n + 255) / 256; return b < kBoxBlocks ? b : kBoxBlocks; }

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
    // up to 2^25 primitives the sort keys are 32 bits (at least 7 bits beyond log2 n: measured as good as 40 on the 20 M-triangle terrain,
    // profiles/r01r_sort_passes.txt): two thirds of the sort's traffic, half its registers and shared memory
    if (!forced_passes && n <= (1u << 25) && passes > 4) passes = 4;
    const bool narrow = passes <= 4;
    if (narrow) k_morton<uint32_t><<<(n + 255) / 256, 256, 0, s>>>(sc.boxes, n, sc.header, reinterpret_cast<uint32_t *>(sc.keys), sc.vals, 63 - 8 * passes);
    else k_morton<uint64_t><<<(n + 255) / 256, 256, 0, s>>>(sc.boxes, n, sc.header, sc.keys, sc.vals, 63 - 8 * passes);
    lc.count++;
    bool in_alt = sort_pairs(s, n, sc.keys, sc.vals, sc.keys_alt, sc.vals_alt, sc.sort_scratch, 0, passes, narrow ? 4 : 8, lc);
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
        k_ploc_finish<<<1, 1, 0, s>>>(a, sc.ploc_state, sc.header, sc.queue); lc.count++;
    } else {
        if (narrow) k_hierarchy<uint32_t><<<(n + 127) / 128, 128, 0, s>>>(reinterpret_cast<const uint32_t *>(keys), vals, sc.boxes, n, sc.bin, sc.flags, sc.header, sc.queue);
        else k_hierarchy<uint64_t><<<(n + 127) / 128, 128, 0, s>>>(keys, vals, sc.boxes, n, sc.bin, sc.flags, sc.header, sc.queue);
        lc.count++;
    }
    // one 8-lane group per wide node in flight; every resident slot when the tree is large enough (the queue needs no co-residency)
    auto launch = [&](auto kernel) {
        static int max_blocks = 0;
        if (!max_blocks) {
            int dev = 0, sms = 0, per_sm = 0;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kCollapseThreads, 0);
            max_blocks = sms * (per_sm > 0 ? per_sm : 1);
        }
        uint32_t blocks = (n / 6 + kCollapseGroups - 1) / kCollapseGroups + 1;
        if (blocks > (uint32_t)max_blocks) blocks = (uint32_t)max_blocks;
        kernel<<<blocks, kCollapseThreads, 0, s>>>(sc.bin, vals, n, sc.header, sc.queue, nodes, capacity, sink); lc.count++;
    };
    if (ploc && n > 1) launch(k_collapse<Sink, false>); else launch(k_collapse<Sink, true>);  // LBVH ids are split positions (contiguous subtrees)
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
    k_init_header<<<init_blocks, 256, 0, s>>>(sc.header, sc.flags, n, sc.queue, (n < capacity ? n : capacity)); lc.count++;
    k_triangle_boxes<<<box_blocks(n), 256, 0, s>>>(in, n, sc.boxes, sc.header); lc.count++;
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
    uint32_t *slot_prim = reinterpret_cast<uint32_t *>(sc.flags);  // the arrival flags are dead once the binary tree exists
    LeafSinkTriangles sink{slot_prim};
    run_pipeline_after_boxes(s, n, sc, nodes, capacity, sink, lc, use_ploc);
    k_pack_tris<<<(n + 255) / 256, 256, 0, s>>>(in, slot_prim, tris, n); lc.count++;
    return use_ploc && n > 1 ? kBuilderPloc : kBuilderLbvh;
}

void build_procedural(cudaStream_t s, uint32_t n, const uint8_t *aabbs, const BuildScratch &sc, WideNode *nodes, uint32_t capacity, PackedTri *slots, LaunchCounter &lc) {
    uint32_t init_blocks = (n + 255) / 256; if (init_blocks > 1024) init_blocks = 1024;
    k_init_header<<<init_blocks, 256, 0, s>>>(sc.header, sc.flags, n, sc.queue, (n < capacity ? n : capacity)); lc.count++;
    k_aabb_boxes<<<box_blocks(n), 256, 0, s>>>(aabbs, n, sc.boxes, sc.header); lc.count++;
    uint32_t *slot_prim = reinterpret_cast<uint32_t *>(sc.flags);
    LeafSinkTriangles sink{slot_prim};
    run_pipeline_after_boxes(s, n, sc, nodes, capacity, sink, lc, false);
    k_pack_aabbs<<<(n + 255) / 256, 256, 0, s>>>(aabbs, slot_prim, slots, n); lc.count++;
}

void build_curves(cudaStream_t s, uint32_t n, const CurveInput &in, const BuildScratch &sc, WideNode *nodes, uint32_t capacity, PackedTri *slots, LaunchCounter &lc) {
    uint32_t init_blocks = (n + 255) / 256; if (init_blocks > 1024) init_blocks = 1024;
    k_init_header<<<init_blocks, 256, 0, s>>>(sc.header, sc.flags, n, sc.queue, (n < capacity ? n : capacity)); lc.count++;
    k_curve_boxes<<<box_blocks(n), 256, 0, s>>>(in, n, sc.boxes, sc.header); lc.count++;
    uint32_t *slot_prim = reinterpret_cast<uint32_t *>(sc.flags);
    LeafSinkTriangles sink{slot_prim};
    run_pipeline_after_boxes(s, n, sc, nodes, capacity, sink, lc, false);
    k_pack_curves<<<(n + 255) / 256, 256, 0, s>>>(in, slot_prim, slots, n); lc.count++;
    if (in.basis != kCurveLinear) { const uint32_t n_segs = n / in.pieces; k_curve_coefs<<<(n_segs + 255) / 256, 256, 0, s>>>(in, slots, n, n_segs); lc.count++; }
}

void build_tlas(cudaStream_t s, uint32_t n, const uint32_t *active_ids, const InstanceRec *instances, const BuildScratch &sc, WideNode *nodes,
                uint32_t *prim_ids, LaunchCounter &lc) {
    uint32_t init_blocks = (n + 255) / 256; if (init_blocks > 1024) init_blocks = 1024;
    k_init_header<<<init_blocks, 256, 0, s>>>(sc.header, sc.flags, n, sc.queue, n); lc.count++;
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

// One 8-lane group per node, lane s = slot s (the layout of k_collapse): every lane fetches its own child — the box of an internal child,
// the one or two triangles of a leaf child — so a node costs one round of loads, not eight in sequence, and nothing lives in local
// memory; node frame by group butterflies, the lane quantises its own slot into the node's copy in shared memory.
constexpr int kRefitThreads = 256;
__global__ void __launch_bounds__(kRefitThreads) k_refit(WideNode *nodes, const PackedTri *tris, uint32_t n_nodes, const uint32_t *__restrict__ parent,
                                               float *boxes, uint32_t *counters) {
    __shared__ WideNode s_node[kRefitThreads / 8];
    const uint32_t lane = threadIdx.x & 31u, sub = lane & 7u;
    WideNode &node = s_node[threadIdx.x >> 3];
    uint32_t i = (blockIdx.x * kRefitThreads + threadIdx.x) >> 3;
    bool live = i < n_nodes && nodes[i].imask == 0;  // nodes with internal children are reached later through them
    while (__any_sync(0xffffffffu, live)) {
        float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
        uint32_t meta = 0;
        if (live) {
            if (sub < (uint32_t)kNodeQuads) reinterpret_cast<uint4 *>(&node)[sub] = reinterpret_cast<const uint4 *>(&nodes[i])[sub];
        }
        __syncwarp();
        if (live) {
            meta = node.meta[sub];
            const uint32_t imask = node.imask;
            if (meta != 0) {
                if (imask >> sub & 1u) {
                    const float *b = boxes + 6 * (size_t)(node.child_base + __popc(imask & ((1u << sub) - 1u)));
                    for (int k = 0; k < 3; k++) { lo[k] = __ldcg(b + k); hi[k] = __ldcg(b + 3 + k); }
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
            }
        }
        float nlo[3], nhi[3], inv_scale[3]; uint32_t ex[3];
#pragma unroll
        for (int k = 0; k < 3; k++) {
            nlo[k] = ordered_to_float(group_min(float_to_ordered(lo[k])));
            nhi[k] = ordered_to_float(group_max(float_to_ordered(hi[k])));
            const float sc = __fdiv_ru(__fsub_ru(nhi[k], nlo[k]), (float)kQMax);
            const uint32_t bits = __float_as_uint(sc);
            uint32_t e = (bits >> 23) + ((bits & 0x7fffffu) ? 1u : 0u);
            e = max(e, 1u); e = min(e, 253u);
            ex[k] = e;
            inv_scale[k] = __uint_as_float((254u - e) << 23);
        }
        uint32_t next = 0xffffffffu;
        if (live) {
            if (meta != 0) for (int k = 0; k < 3; k++) quantise_axis(lo[k], hi[k], nlo[k], inv_scale[k], node.q[k][0][sub], node.q[k][1][sub]);
            if (sub < 3) { node.org[sub] = nlo[sub]; node.e[sub] = (uint8_t)ex[sub]; }
            if (sub < 6) __stcg(boxes + 6 * (size_t)i + sub, sub < 3 ? nlo[sub] : nhi[sub - 3]);
            release_fence();  // the node's box before the arrival below
        }
        __syncwarp();
        if (live) {
            if (sub < (uint32_t)kNodeQuads) reinterpret_cast<uint4 *>(&nodes[i])[sub] = reinterpret_cast<const uint4 *>(&node)[sub];
            if (sub == 0) {
                const uint32_t p = parent[i];
                if (p != 0xffffffffu) {
                    const uint32_t arrived = atomicAdd(&counters[p], 1u) + 1u;
                    if (arrived == (uint32_t)__popc(nodes[p].imask)) next = p;  // the last internal child to arrive takes the parent
                }
            }
        }
        next = GSHFL(next, 0);
        __syncwarp();  // the shared copy is rewritten by the next round
        live = live && next != 0xffffffffu;
        i = next;
    }
}

#undef GSHFL
#undef GXOR

void build_refit_arrays(cudaStream_t s, uint32_t n_nodes, const WideNode *nodes, const RefitArrays &ra, LaunchCounter &lc) {
    if (!n_nodes) return;
    k_refit_parents<<<(n_nodes + 255) / 256, 256, 0, s>>>(nodes, n_nodes, ra.parent, ra.counters); lc.count++;
}

void refit_blas(cudaStream_t s, uint32_t n_nodes, uint32_t n_tris, const TriangleInput &in, WideNode *nodes, PackedTri *tris, const RefitArrays &ra,
                BuildHeader *, LaunchCounter &lc) {
    if (!n_nodes || !n_tris) return;
    k_refit_reset<<<(n_nodes + 255) / 256, 256, 0, s>>>(ra.counters, n_nodes); lc.count++;
    k_pack_tris<<<(n_tris + 255) / 256, 256, 0, s>>>(in, nullptr, tris, n_tris); lc.count++;
    k_refit<<<(unsigned)(((size_t)n_nodes * 8 + kRefitThreads - 1) / kRefitThreads), kRefitThreads, 0, s>>>(nodes, tris, n_nodes, ra.parent, ra.boxes, ra.counters); lc.count++;
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
