// rt_types.cuh — HBM-resident data layouts of the B200 ray-tracing device.
//
// Everything the traversal kernels touch is laid out for full-sector access with sm_100's 256-bit loads
// (LDG.E.ENL2.256: one 32-byte sector per lane and instruction — on divergent addresses the L1TEX data pipe
// charges per lane and instruction, so wider loads are what relieves it; tools/micro/gather256.cu):
//   WideNode     96 B,  32-byte aligned : three sectors, fetched as 3 x LDG.256 (header | x, y planes lo,hi | z planes lo,hi + spare);
//                                         with 16-bit planes (LCB_NODE_BITS=16) 128 B = four sectors (header | x | y | z planes lo,hi)
//   PackedTri    64 B,  64-byte aligned : LDG.256 (v0|prim, v1) + LDG.128 (v2); 16 B spare
//   InstanceRec 128 B,  16-byte aligned : eight float4
#pragma once
#ifdef __CUDACC_RTC__
// NVRTC (IR-lowered kernels, shader.cu): no host headers; the fixed-width names and offsetof come from the compiler
typedef signed char int8_t; typedef unsigned char uint8_t; typedef short int16_t; typedef unsigned short uint16_t;
typedef int int32_t; typedef unsigned int uint32_t; typedef long long int64_t; typedef unsigned long long uint64_t;
typedef unsigned long size_t;
#ifndef INFINITY
#define INFINITY __int_as_float(0x7f800000)
#endif
#else
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <cuda_runtime.h>
#endif

namespace lcb {

// 8-wide BVH node with 16-bit quantised child planes (a compressed-wide-BVH in the spirit of
// Ylitie, Karras, Laine 2017, widened so that a node is exactly one 128-byte line).
//   child plane k of child i :  org[k] + q[k][i] * 2^(e[k]-127)
//   meta[i] : 0 = empty slot
//             internal child : 0x20 | (24 + i)                      (i = slot)
//             leaf child     : (unary count 1|3|7) << 5 | first prim offset (0..23)
//   imask   : bit i set <=> slot i is an internal child; internal children are stored
//             contiguously from child_base in ascending slot order
//   prim_base : first PackedTri (BLAS) / first entry of the instance-id list (TLAS)
// Plane resolution (compile-time): 8 bits -> a 96-byte node = THREE 32-byte sectors per visit, 16 bits -> 128 bytes = four.  The traversal
// is bound by the L1TEX wavefronts of exactly these divergent sector fetches, so the narrower node is the default; its boxes are looser by
// at most 1/255 of the parent's extent per plane (a few per cent more node visits), which the sector count more than pays for
// (profiles/r02g_node_bits.txt).  -DLCB_NODE_BITS=16 builds the round-1 layout for A/B runs.
#ifndef LCB_NODE_BITS
#define LCB_NODE_BITS 8
#endif
#if LCB_NODE_BITS == 16
typedef uint16_t qplane_t;
#define LCB_NODE_ALIGN 128
#else
typedef uint8_t qplane_t;
#define LCB_NODE_ALIGN 32
#endif
constexpr uint32_t kQMax = (1u << LCB_NODE_BITS) - 1u;  // largest plane index
struct alignas(LCB_NODE_ALIGN) WideNode {
    float org[3];
    uint8_t e[3];
    uint8_t imask;
    uint32_t child_base;
    uint32_t prim_base;
    uint8_t meta[8];
    qplane_t q[3][2][8];  // [axis][0 = lower plane, 1 = upper plane][slot]
#if LCB_NODE_BITS == 8
    uint8_t spare[16];    // the node is three whole sectors: header | x, y planes | z planes + spare
#endif
};
static_assert(sizeof(WideNode) == (LCB_NODE_BITS == 16 ? 128 : 96), "WideNode is a whole number of 32-byte sectors");
constexpr int kNodeQuads = sizeof(WideNode) / 16;  // 16-byte pieces of a node (node stores)

struct alignas(64) PackedTri {
    float v0[3]; uint32_t prim;
    float v1[3]; uint32_t pad1;
    float v2[3]; uint32_t pad2;
    uint32_t spare[4];
};
static_assert(sizeof(PackedTri) == 64, "PackedTri is 64 bytes");

// One slot of the instance table (what AccelImpl keeps per instance, cpu/accel.rs:270-296).
struct alignas(16) InstanceRec {
    float inv[12];            // world -> object, row-major 3x4
    const WideNode *nodes;    // BLAS nodes (nullptr: empty mesh or invalid slot)
    const PackedTri *tris;    // BLAS packed triangles
    uint32_t visibility;
    uint32_t user_id;
    uint32_t flags;           // bit0 valid, bit1 opaque, bit2 procedural primitive, bit3 curve, bit4 `inv` is exactly the identity (trace_device.cuh enter_instance)
    uint32_t pad;
    float affine[12];         // object -> world as given by the frontend (for instance_transform)
};
static_assert(sizeof(InstanceRec) == 128, "InstanceRec is 128 bytes");

// Binary LBVH node produced by the fused hierarchy+refit kernel, consumed by the collapse.
// child id: bit31 set -> leaf, low bits = position in the sorted primitive order.
// Each child owns two float4 so the two arriving threads never write the same 16 bytes.
struct alignas(16) BinNode {
    float llo[3]; uint32_t left;
    float lhi[3]; uint32_t lcount;
    float rlo[3]; uint32_t right;
    float rhi[3]; uint32_t rcount;
};
static_assert(sizeof(BinNode) == 64, "BinNode is 64 bytes");

struct alignas(16) PrimBox { float lo[3]; uint32_t pad0; float hi[3]; uint32_t pad1; };

// device-side scratch header of one build
struct BuildHeader {
    int bounds_lo[3];   // ordered-int encoded float min of centroids
    int bounds_hi[3];
    uint32_t root;          // binary root id
    uint32_t emitted;       // primitives emitted into leaves (= prim_count after a complete collapse)
    uint32_t node_count;    // wide nodes allocated   } one 8-byte aligned pair: the collapse allocates both
    uint32_t prim_count;    // packed prims allocated } with a single 64-bit atomic per CTA and step
    uint32_t tickets;       // collapse: next wide-node id to hand out as a work ticket
    uint32_t max_depth;
    uint32_t error;         // nonzero: builder failure code
    uint32_t unused0;
    float root_lo[3]; float prim_area_sum;  // sum of the primitives' box half-areas (builder choice, bvh_build.cu)
    float root_hi[3]; float pad2;
    uint32_t pad_line[24];  // keeps collapse_done out of the cache line of the allocation atomics
    uint32_t collapse_done; // collapse: final wide-node count + 1 once every primitive has a leaf slot (1 after an error); polled by idle warps
    uint32_t reserved[23];
};
#ifndef __CUDACC_RTC__
static_assert(offsetof(BuildHeader, node_count) % 8 == 0, "node_count/prim_count must form an aligned 64-bit word");
static_assert(offsetof(BuildHeader, collapse_done) / 128 != offsetof(BuildHeader, tickets) / 128 && offsetof(BuildHeader, collapse_done) / 128 != offsetof(BuildHeader, node_count) / 128, "collapse_done has its own cache line");
#endif

// What a traversal needs to know about one acceleration structure (by value in kernel parameters).
struct AccelView {
    const WideNode *tlas_nodes;      // nullptr => empty accel
    const uint32_t *tlas_prims;      // instance ids referenced by TLAS leaves
    const InstanceRec *instances;
    uint32_t instance_count;
    float world_lo[3], world_hi[3];  // bounds of the TLAS root (ray reordering quantises origins against them)
    uint32_t flags;                  // bit 0: some instance is a curve (the batch entry points then take the per-thread kernel)
};
static_assert(sizeof(AccelView) == 56, "AccelView is mirrored in lc_accel / HostAccelArg");

constexpr uint32_t kCurveSubdiv = 8;  // pieces per cubic curve segment (trace_device.cuh "curves")
constexpr int kMaxWideDepth = 40;   // builder fails loudly beyond this; traversal stack is sized for it
// Entries of a traversal stack.  Mandatory pushes: one node group per level + three per instance entry (<= 2 * kMaxWideDepth + 3 = 83);
// the if-if loop may also postpone one primitive group per level, and stops doing so at kPostponeLimit (trace_device.cuh) so that the two
// together never exceed the stack.
constexpr int kTraversalStack = 128;

__host__ __device__ inline int float_to_ordered(float f) {
    int i;
#ifdef __CUDA_ARCH__
    i = __float_as_int(f);
#else
    memcpy(&i, &f, 4);
#endif
    return i >= 0 ? i : i ^ 0x7fffffff;
}
__host__ __device__ inline float ordered_to_float(int i) {
    i = i >= 0 ? i : i ^ 0x7fffffff;
#ifdef __CUDA_ARCH__
    return __int_as_float(i);
#else
    float f; memcpy(&f, &i, 4); return f;
#endif
}

}  // namespace lcb
