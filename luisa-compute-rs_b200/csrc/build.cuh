// build.cuh — internal interface between the host-side device object (device.cu) and the
// kernel translation units (radix_sort.cu, bvh_build.cu, trace.cu).
#pragma once
#include "rt_types.cuh"
#include <atomic>
#include <cstddef>

namespace lcb {

struct LaunchCounter { std::atomic<unsigned long long> count{0}; };  // streams of one device dispatch from several host threads

// ---- radix_sort.cu -------------------------------------------------------------------------
size_t sort_scratch_bytes(uint32_t n, int passes);
bool sort_pairs(cudaStream_t s, uint32_t n, void *keys, uint32_t *vals, void *keys_alt, uint32_t *vals_alt, void *scratch,
                int begin_bit, int passes, int key_bytes /* 4: uint32_t keys, 8: uint64_t */, LaunchCounter &lc);

// ---- bvh_build.cu --------------------------------------------------------------------------
struct TriangleInput {
    const uint8_t *vertices; size_t vertex_stride;
    const uint8_t *indices;  // 12-byte uint3 records
};

struct PlocState { uint32_t n_clusters; uint32_t n_nodes; uint32_t iterations; uint32_t pad; };

// Scratch carved out of one allocation; sizes from build_scratch_layout().
struct BuildScratch {
    BuildHeader *header;
    PrimBox *boxes;          // n
    uint64_t *keys, *keys_alt;
    uint32_t *vals, *vals_alt;
    void *sort_scratch;
    BinNode *bin;            // n-1
    int *flags;              // n-1
    unsigned long long *queue;  // n (collapse work items)
    float4 *ploc_a, *ploc_b;    // PLOC cluster arrays (2 float4 per cluster), double-buffered
    uint32_t *ploc_counts;      // per-tile survivor counts / offsets
    PlocState *ploc_state;
    size_t total_bytes;
};
BuildScratch build_scratch_layout(void *base, uint32_t n);

// Full build: prim boxes -> Morton -> sort -> fused hierarchy+refit -> collapse to WideNode + leaves.
// `nodes` has room for `capacity` wide nodes (the builder raises its error flag beyond that); `tris` capacity `n` (BLAS) / `prim_ids` capacity n (TLAS).
// builder: how the binary tree over the Morton-sorted primitives is formed — the LBVH split rule (k_hierarchy), PLOC (agglomerative
// clustering, k_ploc_*), or chosen per mesh from the primitives' overlap (kBuilderAuto).
enum { kBuilderLbvh = 0, kBuilderPloc = 1, kBuilderAuto = 2 };
// returns the builder that ran (kBuilderLbvh / kBuilderPloc)
int build_blas(cudaStream_t s, uint32_t n_tris, const TriangleInput &in, const BuildScratch &sc, WideNode *nodes, uint32_t capacity, PackedTri *tris, LaunchCounter &lc, int builder);
// Procedural primitives: a BLAS over user AABBs (24-byte {min, max} records); leaf slots hold the box and the primitive id.
void build_procedural(cudaStream_t s, uint32_t n, const uint8_t *aabbs, const BuildScratch &sc, WideNode *nodes, uint32_t capacity, PackedTri *slots, LaunchCounter &lc);
// Curves: a BLAS over the rounded-cone pieces of the segments (trace_device.cuh "curves"); n = seg_count * pieces per segment.
struct CurveInput { const uint8_t *cps; size_t cp_stride; const uint32_t *segs; uint32_t basis; uint32_t pieces; };
void build_curves(cudaStream_t s, uint32_t n, const CurveInput &in, const BuildScratch &sc, WideNode *nodes, uint32_t capacity, PackedTri *slots, LaunchCounter &lc);
void build_tlas(cudaStream_t s, uint32_t n_active, const uint32_t *active_ids, const InstanceRec *instances, const BuildScratch &sc,
                WideNode *nodes, uint32_t *prim_ids, LaunchCounter &lc);

// Refit (PreferUpdate): rewrites packed triangles from the current vertex data and recomputes
// every node's quantised child planes bottom-up.  `parent` / `node_boxes` are side arrays kept
// from the build.
struct RefitArrays {
    uint32_t *parent;      // per wide node: parent index (root: 0xffffffff)
    float *boxes;          // per wide node: 6 floats, full-precision bounds
    uint32_t *counters;    // per wide node: arrival counter
};
void refit_blas(cudaStream_t s, uint32_t n_nodes, uint32_t n_tris, const TriangleInput &in, WideNode *nodes, PackedTri *tris,
                const RefitArrays &ra, BuildHeader *header, LaunchCounter &lc);
void build_refit_arrays(cudaStream_t s, uint32_t n_nodes, const WideNode *nodes, const RefitArrays &ra, LaunchCounter &lc);

// Instance table maintenance (AccelImpl::update, cpu/accel.rs:354-427): expanded modification
// records are scattered into the device table in one launch.
// Each record is the complete new state of one slot (the host mirror resolves the reference's
// flag-ordering rules and sends one record per touched slot).
struct InstanceModRec {
    uint32_t index, flags, visibility, user_id;
    const WideNode *nodes; const PackedTri *tris;
    float affine[12];
    float inv[12];
};
void apply_instance_mods(cudaStream_t s, InstanceRec *table, const InstanceModRec *mods, uint32_t n_mods, LaunchCounter &lc);

// ---- trace.cu ------------------------------------------------------------------------------
struct TraceCounters { unsigned long long nodes_visited, tris_tested, rays, instance_entries; };

// AccelView: rt_types.cuh (shared with NVRTC-compiled kernels)

// Candidate hook of the batch RayQuery entry points (evaluated on device for triangles of non-opaque instances).
struct CandidateFilter {
    int kind;                  // 0 commit all, 1 barycentric disc (examples/ray_query.rs), 2 per-primitive cut-out bits, 3 reject all
    float radius;              // kind 1
    const uint32_t *bits;      // kind 2: bit (first_bit[inst] + prim) set => commit
    const uint32_t *first_bit; // kind 2: one entry per instance slot
};

void trace_closest(cudaStream_t s, const AccelView &a, const void *rays, void *hits, uint64_t count, uint32_t mask, unsigned long long *work_counter,
                   TraceCounters *counters /* device, nullable */, LaunchCounter &lc, unsigned grid_limit = 0 /* CTAs; 0 = every resident slot */);
void trace_any(cudaStream_t s, const AccelView &a, const void *rays, uint32_t *occluded, uint64_t count, uint32_t mask, unsigned long long *work_counter,
               LaunchCounter &lc, unsigned grid_limit = 0);

void ray_query(cudaStream_t s, const AccelView &a, const void *rays, void *committed_hits, uint64_t count, uint32_t mask, bool terminate_on_first,
               const CandidateFilter &filter, unsigned long long *work_counter, LaunchCounter &lc);

// ---- path_tracer.cu ------------------------------------------------------------------------
void launch_path_tracer(cudaStream_t s, const AccelView &accel, const float *const *vertex_heap, const uint32_t *const *index_heap, float4 *image,
                        uint32_t *seed_image, uint32_t width, uint32_t height, uint32_t spp_per_dispatch, uint32_t max_depth, float tan_half_fov,
                        unsigned long long *ray_counters, LaunchCounter &lc);

}  // namespace lcb
