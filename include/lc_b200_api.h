/*
 * lc_b200_api.h — C ABI of the B200 ray-tracing device for luisa-compute-rs.
 *
 * Two groups of declarations:
 *
 *  (1) A layout-compatible mirror of the reference's FFI contract
 *      (luisa_compute_api_types, cbindgen output
 *      LC/include/luisa/rust/api_types.h).  The Rust frontend binds the single
 *      symbol `luisa_compute_lib_interface` (backend/lib.rs:32-41) and from then
 *      on only calls through the two function-pointer tables below, so field
 *      ORDER, SIZE and ALIGNMENT are the contract; type names are ours
 *      (prefix lcb_) and every struct carries a static assertion of the size
 *      the reference header produces.  Only the parts of the contract that the
 *      hot path (MeshBuild / AccelBuild / ray queries and the buffer plumbing
 *      around them) needs are given behaviour; every other table slot is
 *      filled with a function that logs "unsupported" and aborts, as the
 *      reference does for failures (backend_impl/src/lib.rs:101-131).
 *
 *  (2) Native batch entry points (`lc_b200_*`): the same operations the
 *      reference hands to JIT-ed kernels one ray at a time through the
 *      `defs::Accel` vtable (luisa_compute_cpu_kernel_defs/src/lib.rs:112-125:
 *      trace_closest / trace_any / instance_* accessors), here applied to a
 *      whole device- or host-resident buffer of rays.
 *
 * Plain C, no CUDA or torch types in any signature.
 */
#ifndef LC_B200_API_H
#define LC_B200_API_H

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#define LCB_STATIC_ASSERT(c, m) static_assert(c, m)
#else
#define LCB_STATIC_ASSERT(c, m) _Static_assert(c, m)
#endif

#if defined(_WIN32)
#define LCB_EXPORT __declspec(dllexport)
#else
#define LCB_EXPORT __attribute__((visibility("default")))
#endif

#define LCB_INVALID_HANDLE UINT64_MAX /* api_types/src/lib.rs:5 */

/* ------------------------------------------------------------------------- */
/* Ray-tracing value types (byte layouts: SURVEY.md appendix B)               */
/* ------------------------------------------------------------------------- */

/* rtx.rs:329-338, cpu_kernel_defs/src/lib.rs:42-54 — 32 B, align 16 */
typedef struct __attribute__((aligned(16))) lcb_ray {
    float orig[3];
    float tmin;
    float dir[3];
    float tmax;
} lcb_ray;
LCB_STATIC_ASSERT(sizeof(lcb_ray) == 32, "Ray is 32 bytes");

/* rtx.rs:356-366 (SurfaceHit) == cpu_kernel_defs TriangleHit/Hit :94-105 — 24 B, align 8.
 * miss: inst = prim = 0xffffffff, bary = (0,0), t = ray.tmax (cpu/accel.rs:502-507). */
typedef struct __attribute__((aligned(8))) lcb_surface_hit {
    uint32_t inst;
    uint32_t prim;
    float bary[2]; /* P = (1-u-v) v0 + u v1 + v v2  (rtx.rs:384) */
    float committed_ray_t;
    uint32_t _pad;
} lcb_surface_hit;
LCB_STATIC_ASSERT(sizeof(lcb_surface_hit) == 24, "SurfaceHit is 24 bytes");

/* rtx.rs:477-485, cpu_kernel_defs :68-77 */
typedef struct __attribute__((aligned(8))) lcb_committed_hit {
    uint32_t inst;
    uint32_t prim;
    float bary[2];
    uint32_t hit_type; /* 0 miss, 1 triangle, 2 procedural */
    float committed_ray_t;
} lcb_committed_hit;
LCB_STATIC_ASSERT(sizeof(lcb_committed_hit) == 24, "CommittedHit is 24 bytes");

typedef struct lcb_aabb { float min[3]; float max[3]; } lcb_aabb;      /* rtx.rs:339-345 */
typedef struct lcb_triangle { uint32_t i[3]; } lcb_triangle;           /* rtx.rs:536 (Index) */
LCB_STATIC_ASSERT(sizeof(lcb_triangle) == 12, "Index is 12 bytes");

/* ------------------------------------------------------------------------- */
/* Handles, options and enums (api_types/src/lib.rs:188-253)                  */
/* ------------------------------------------------------------------------- */

typedef struct lcb_handle { uint64_t id; } lcb_handle;
typedef lcb_handle lcb_context, lcb_device, lcb_buffer, lcb_texture, lcb_stream, lcb_event,
    lcb_shader, lcb_swapchain, lcb_bindless, lcb_mesh, lcb_curve, lcb_procedural, lcb_accel;

enum { LCB_REQUEST_PREFER_UPDATE = 0, LCB_REQUEST_FORCE_BUILD = 1 };  /* AccelBuildRequest */
enum { LCB_HINT_FAST_TRACE = 0, LCB_HINT_FAST_BUILD = 1 };            /* AccelUsageHint   */
enum { LCB_STREAM_GRAPHICS = 0, LCB_STREAM_COMPUTE = 1, LCB_STREAM_COPY = 2 };

typedef struct lcb_accel_option { /* api_types:204-220 — 8 B */
    int32_t hint;
    bool allow_compaction;
    bool allow_update;
} lcb_accel_option;
LCB_STATIC_ASSERT(sizeof(lcb_accel_option) == 8, "AccelOption is 8 bytes");

/* AccelBuildModificationFlags bits, api_types:222-236 */
enum {
    LCB_MOD_PRIMITIVE = 1u << 0,
    LCB_MOD_TRANSFORM = 1u << 1,
    LCB_MOD_OPAQUE_ON = 1u << 2,
    LCB_MOD_OPAQUE_OFF = 1u << 3,
    LCB_MOD_VISIBILITY = 1u << 4,
    LCB_MOD_USER_ID = 1u << 5
};

typedef struct lcb_accel_modification { /* api_types:244-253 — 72 B */
    uint32_t index;
    uint32_t user_id;
    uint32_t flags;
    uint32_t visibility;
    uint64_t mesh;    /* mesh / curve / procedural handle */
    float affine[12]; /* row-major 3x4 */
} lcb_accel_modification;
LCB_STATIC_ASSERT(sizeof(lcb_accel_modification) == 72, "AccelBuildModification is 72 bytes");

/* ------------------------------------------------------------------------- */
/* Commands (api_types:486-732).  Tag values = declaration order.             */
/* ------------------------------------------------------------------------- */

enum {
    LCB_CMD_BUFFER_UPLOAD = 0,
    LCB_CMD_BUFFER_DOWNLOAD = 1,
    LCB_CMD_BUFFER_COPY = 2,
    LCB_CMD_BUFFER_TO_TEXTURE = 3,
    LCB_CMD_TEXTURE_TO_BUFFER = 4,
    LCB_CMD_TEXTURE_UPLOAD = 5,
    LCB_CMD_TEXTURE_DOWNLOAD = 6,
    LCB_CMD_TEXTURE_COPY = 7,
    LCB_CMD_SHADER_DISPATCH = 8,
    LCB_CMD_MESH_BUILD = 9,
    LCB_CMD_CURVE_BUILD = 10,
    LCB_CMD_PROCEDURAL_BUILD = 11,
    LCB_CMD_ACCEL_BUILD = 12,
    LCB_CMD_BINDLESS_UPDATE = 13
};

typedef struct lcb_cmd_buffer_upload { lcb_buffer buffer; size_t offset, size; const uint8_t *data; } lcb_cmd_buffer_upload;
typedef struct lcb_cmd_buffer_download { lcb_buffer buffer; size_t offset, size; uint8_t *data; } lcb_cmd_buffer_download;
typedef struct lcb_cmd_buffer_copy { lcb_buffer src; size_t src_offset; lcb_buffer dst; size_t dst_offset, size; } lcb_cmd_buffer_copy;
typedef struct lcb_cmd_buffer_texture { lcb_buffer buffer; size_t buffer_offset; lcb_texture texture; int32_t storage; uint32_t level; uint32_t size[3]; } lcb_cmd_buffer_texture;
typedef struct lcb_cmd_texture_transfer { lcb_texture texture; int32_t storage; uint32_t level; uint32_t size[3]; uint8_t *data; } lcb_cmd_texture_transfer;
typedef struct lcb_cmd_texture_copy { int32_t storage; lcb_texture src, dst; uint32_t size[3]; uint32_t src_level, dst_level; } lcb_cmd_texture_copy;

enum { LCB_ARG_BUFFER = 0, LCB_ARG_TEXTURE = 1, LCB_ARG_UNIFORM = 2, LCB_ARG_BINDLESS = 3, LCB_ARG_ACCEL = 4 };
typedef struct lcb_argument { /* api_types:454-484 — 32 B */
    int32_t tag;
    union {
        struct { lcb_buffer buffer; size_t offset, size; } buffer;
        struct { lcb_texture texture; uint32_t level; } texture;
        struct { const uint8_t *data; size_t size; } uniform;
        lcb_bindless bindless;
        lcb_accel accel;
    } u;
} lcb_argument;
LCB_STATIC_ASSERT(sizeof(lcb_argument) == 32, "Argument is 32 bytes");

typedef struct lcb_cmd_shader_dispatch { lcb_shader shader; uint32_t dispatch_size[3]; const lcb_argument *args; size_t args_count; } lcb_cmd_shader_dispatch;

typedef struct lcb_cmd_mesh_build { /* api_types:603-616 */
    lcb_mesh mesh;
    int32_t request;
    lcb_buffer vertex_buffer;
    size_t vertex_buffer_offset, vertex_buffer_size, vertex_stride;
    lcb_buffer index_buffer;
    size_t index_buffer_offset, index_buffer_size, index_stride; /* must be 12 (api/runtime.cpp:191) */
} lcb_cmd_mesh_build;
LCB_STATIC_ASSERT(sizeof(lcb_cmd_mesh_build) == 80, "MeshBuildCommand is 80 bytes");

/* CurveBuildCommand (api_types:618-631; backend: GeometryImpl::build_curve, cpu/accel.rs:142-203).  basis = api::CurveBasis
 * (0 PiecewiseLinear, 1 CubicBSpline, 2 CatmullRom, 3 Bezier); control points are float4 {x, y, z, radius} at cp_stride (>= 16,
 * multiple of 16); segments are u32 indices of each segment's first control point.  Always a full build. */
typedef struct lcb_cmd_curve_build { lcb_curve curve; int32_t request; int32_t basis; size_t cp_count, seg_count; lcb_buffer cp_buffer; size_t cp_offset, cp_stride; lcb_buffer seg_buffer; size_t seg_offset; } lcb_cmd_curve_build;
typedef struct lcb_cmd_procedural_build { lcb_procedural handle; int32_t request; lcb_buffer aabb_buffer; size_t aabb_offset, aabb_count; } lcb_cmd_procedural_build;

typedef struct lcb_cmd_accel_build { /* api_types:643-652 */
    lcb_accel accel;
    int32_t request;
    uint32_t instance_count;
    const lcb_accel_modification *modifications;
    size_t modifications_count;
    bool update_instance_buffer_only;
} lcb_cmd_accel_build;
LCB_STATIC_ASSERT(sizeof(lcb_cmd_accel_build) == 40, "AccelBuildCommand is 40 bytes");

typedef struct lcb_sampler { int32_t filter, address; } lcb_sampler;
typedef struct lcb_bindless_buffer_update { int32_t op; lcb_buffer handle; size_t offset; } lcb_bindless_buffer_update;
typedef struct lcb_bindless_texture_update { int32_t op; lcb_texture handle; lcb_sampler sampler; } lcb_bindless_texture_update;
typedef struct lcb_bindless_modification { size_t slot; lcb_bindless_buffer_update buffer; lcb_bindless_texture_update tex2d, tex3d; } lcb_bindless_modification;
LCB_STATIC_ASSERT(sizeof(lcb_bindless_modification) == 80, "BindlessArrayUpdateModification is 80 bytes");
typedef struct lcb_cmd_bindless_update { lcb_bindless handle; const lcb_bindless_modification *modifications; size_t modifications_count; } lcb_cmd_bindless_update;

typedef struct lcb_command { /* api_types:715-732 — tag + 8-aligned union, 88 B */
    int32_t tag;
    union {
        lcb_cmd_buffer_upload buffer_upload;
        lcb_cmd_buffer_download buffer_download;
        lcb_cmd_buffer_copy buffer_copy;
        lcb_cmd_buffer_texture buffer_to_texture, texture_to_buffer;
        lcb_cmd_texture_transfer texture_upload, texture_download;
        lcb_cmd_texture_copy texture_copy;
        lcb_cmd_shader_dispatch shader_dispatch;
        lcb_cmd_mesh_build mesh_build;
        lcb_cmd_curve_build curve_build;
        lcb_cmd_procedural_build procedural_build;
        lcb_cmd_accel_build accel_build;
        lcb_cmd_bindless_update bindless_update;
    } u;
} lcb_command;
LCB_STATIC_ASSERT(sizeof(lcb_command) == 88, "Command is 88 bytes");
LCB_STATIC_ASSERT(offsetof(lcb_command, u) == 8, "Command payload at offset 8");

typedef struct lcb_command_list { const lcb_command *commands; size_t commands_count; } lcb_command_list;

/* ------------------------------------------------------------------------- */
/* Creation results, shader/swapchain descriptors, extension tables           */
/* ------------------------------------------------------------------------- */

typedef struct lcb_created { uint64_t handle; void *native_handle; } lcb_created;
typedef struct lcb_created_buffer { lcb_created resource; size_t element_stride, total_size_bytes; } lcb_created_buffer;
typedef struct lcb_created_shader { lcb_created resource; uint32_t block_size[3]; } lcb_created_shader;
typedef struct lcb_created_swapchain { lcb_created resource; int32_t storage; } lcb_created_swapchain;
typedef struct lcb_kernel_module { uint64_t ptr; /* *const ir::KernelModule, proxy.rs:188-193 */ } lcb_kernel_module;
typedef struct lcb_shader_option { bool enable_cache, enable_fast_math, enable_debug_info, compile_only, time_trace; uint32_t max_registers; const char *name; const char *native_include; } lcb_shader_option;
typedef struct lcb_swapchain_option { uint64_t display, window; uint32_t width, height; bool wants_hdr, wants_vsync; uint32_t back_buffer_count; } lcb_swapchain_option;
typedef struct lcb_pinned_memory_ext { void *data; void *pin_host_memory; void *allocate_pinned_memory; } lcb_pinned_memory_ext;
typedef struct lcb_denoiser_ext { void *data; void *create; void *init; void *execute; void *destroy; } lcb_denoiser_ext;
typedef void (*lcb_dispatch_callback)(uint8_t *);

typedef struct lcb_logger_message { const char *target, *level, *message; } lcb_logger_message;

/* api_types:773-826 — `device` followed by 35 function pointers, in this order. */
typedef struct lcb_device_interface lcb_device_interface;
struct lcb_device_interface {
    lcb_device device;
    void (*destroy_device)(lcb_device_interface);
    lcb_created_buffer (*create_buffer)(lcb_device, const void *ir_type /* &CArc<ir::Type> */, size_t count, void *ext_mem);
    void (*destroy_buffer)(lcb_device, lcb_buffer);
    lcb_created (*create_texture)(lcb_device, int32_t format, uint32_t dim, uint32_t w, uint32_t h, uint32_t d, uint32_t mips, bool simultaneous, bool raster);
    void *(*native_handle)(lcb_device);
    uint32_t (*compute_warp_size)(lcb_device);
    void (*destroy_texture)(lcb_device, lcb_texture);
    lcb_created (*create_bindless_array)(lcb_device, size_t);
    void (*destroy_bindless_array)(lcb_device, lcb_bindless);
    lcb_created (*create_stream)(lcb_device, int32_t tag);
    void (*destroy_stream)(lcb_device, lcb_stream);
    void (*synchronize_stream)(lcb_device, lcb_stream);
    void (*dispatch)(lcb_device, lcb_stream, lcb_command_list, lcb_dispatch_callback, uint8_t *ctx);
    lcb_created_swapchain (*create_swapchain)(lcb_device, const lcb_swapchain_option *, lcb_stream);
    void (*present_display_in_stream)(lcb_device, lcb_stream, lcb_swapchain, lcb_texture);
    void (*destroy_swapchain)(lcb_device, lcb_swapchain);
    lcb_created_shader (*create_shader)(lcb_device, lcb_kernel_module, const lcb_shader_option *);
    void (*destroy_shader)(lcb_device, lcb_shader);
    lcb_created (*create_event)(lcb_device);
    void (*destroy_event)(lcb_device, lcb_event);
    void (*signal_event)(lcb_device, lcb_event, lcb_stream, uint64_t);
    void (*synchronize_event)(lcb_device, lcb_event, uint64_t);
    void (*wait_event)(lcb_device, lcb_event, lcb_stream, uint64_t);
    bool (*is_event_completed)(lcb_device, lcb_event, uint64_t);
    lcb_created (*create_mesh)(lcb_device, const lcb_accel_option *);
    void (*destroy_mesh)(lcb_device, lcb_mesh);
    lcb_created (*create_curve)(lcb_device, const lcb_accel_option *);
    void (*destroy_curve)(lcb_device, lcb_curve);
    lcb_created (*create_procedural_primitive)(lcb_device, const lcb_accel_option *);
    void (*destroy_procedural_primitive)(lcb_device, lcb_procedural);
    lcb_created (*create_accel)(lcb_device, const lcb_accel_option *);
    void (*destroy_accel)(lcb_device, lcb_accel);
    char *(*query)(lcb_device, const char *);
    lcb_pinned_memory_ext (*pinned_memory_ext)(lcb_device);
    lcb_denoiser_ext (*denoiser_ext)(lcb_device);
};
LCB_STATIC_ASSERT(sizeof(lcb_device_interface) == 8 + 35 * sizeof(void *), "DeviceInterface = device + 35 fn ptrs");

/* api_types:761-771 */
typedef struct lcb_lib_interface {
    void *inner;
    void (*set_logger_callback)(void (*)(lcb_logger_message));
    lcb_context (*create_context)(const char *runtime_dir);
    void (*destroy_context)(lcb_context);
    lcb_device_interface (*create_device)(lcb_context, const char *name, const char *json_config);
    void (*free_string)(char *);
} lcb_lib_interface;

/* THE drop-in symbol.  Replaces LC/src/api/runtime.cpp:828-837 and
 * luisa_compute_backend_impl/src/lib.rs:233-243; bound by backend/lib.rs:32-41.
 * Device names served: "b200" (and "cuda-b200").  Anything else -> logged error + abort. */
LCB_EXPORT lcb_lib_interface luisa_compute_lib_interface(void);

/* ------------------------------------------------------------------------- */
/* Native batch entry points                                                  */
/* ------------------------------------------------------------------------- */

/* Batch form of defs::Accel::trace_closest (cpu_kernel_defs/src/lib.rs:115,
 * cpu/accel.rs:449-509, called per ray from cpu_resource.h:288): for i < count,
 * hits[i] = closest hit of rays[i] against `accel` restricted to instances with
 * (mask & visibility) != 0.  `rays`/`hits` are buffers created by this device;
 * offsets in bytes.  Enqueued on `stream`; returns immediately. */
LCB_EXPORT void lc_b200_trace_closest(lcb_device, lcb_stream, lcb_accel, lcb_buffer rays, size_t rays_offset,
                                      lcb_buffer hits, size_t hits_offset, uint64_t count, uint32_t mask);

/* Batch form of defs::Accel::trace_any (cpu/accel.rs:511-535): out[i] = 1u if any
 * triangle is hit in (tmin, tmax], else 0u; `occluded` holds count uint32. */
LCB_EXPORT void lc_b200_trace_any(lcb_device, lcb_stream, lcb_accel, lcb_buffer rays, size_t rays_offset,
                                  lcb_buffer occluded, size_t occluded_offset, uint64_t count, uint32_t mask);

/* Batch form of RayQuery (defs::Accel::ray_query, cpu_kernel_defs/src/lib.rs:124; AccelImpl::ray_query,
 * cpu/accel.rs:582-800; frontend RayQueryBase rtx.rs:672-756): for i < count, committed[i] is the CommittedHit of
 * `accel.traverse(rays[i], mask)` (terminate_on_first = false, IR RayTracingQueryAll) or `traverse_any` (true,
 * RayTracingQueryAny).  Triangles of opaque instances commit without a callback; each triangle candidate of a
 * NON-opaque instance (AccelBuildModification OPAQUE_OFF) is handed to the candidate hook, which stands in for the
 * DSL's on_surface_hit closure until kernels are lowered from IR: a device-side pure function selected by `filter`.
 * A query that commits nothing returns {inst = prim = ~0, bary = 0, hit_type = 0 (Miss), t = 0}. */
enum {
    LCB_FILTER_COMMIT_ALL = 0, /* candidate.commit() unconditionally */
    LCB_FILTER_BARY_DISC = 1,  /* examples/ray_query.rs:148-162: commit iff |uvw.xy|, |uvw.yz|, |uvw.xz| < radius */
    LCB_FILTER_PRIM_BITS = 2,  /* commit iff bit (first_bit[inst] + prim) of `bits` is set (cut-out table) */
    LCB_FILTER_REJECT_ALL = 3  /* never commit: non-opaque instances are invisible */
};
typedef struct lcb_candidate_filter {
    int32_t kind;
    float radius;        /* BARY_DISC */
    lcb_buffer bits;     /* PRIM_BITS: uint32 words */
    lcb_buffer first_bit; /* PRIM_BITS: one uint32 per instance slot */
} lcb_candidate_filter;
LCB_EXPORT void lc_b200_ray_query(lcb_device, lcb_stream, lcb_accel, lcb_buffer rays, size_t rays_offset, lcb_buffer committed, size_t committed_offset,
                                  uint64_t count, uint32_t mask, bool terminate_on_first, const lcb_candidate_filter *filter);

/* The kernel of luisa_compute/examples/path_tracer.rs:247-455 (Cornell box path tracer: LCG sampler, area-light MIS,
 * cosine-weighted bounces, Russian roulette) in the form the IR -> CUDA lowering will emit: one thread per pixel calling
 * the single-ray traversal routines.  One call = one `path_tracer.dispatch([w, h, 1], &acc_img, &seed_img, &accel, &res)`
 * (path_tracer.rs:541-548): `image` (w*h float4, accumulated radiance + sample count) and `seed_image` (w*h uint32) are
 * read and updated.  vertex_heap[i] / index_heap[i] are the [f32;3] vertex and Index buffers of instance i (the
 * example's bindless heaps).  If ray_counts_out is non-null the call synchronises and returns the number of closest-hit
 * and any-hit rays traced by this dispatch. */
typedef struct lcb_path_tracer_args {
    lcb_accel accel;
    const lcb_buffer *vertex_heap;
    const lcb_buffer *index_heap;
    uint32_t heap_size;
    lcb_buffer image, seed_image;
    uint32_t width, height, spp_per_dispatch, max_depth; /* example: 32 spp per dispatch, depth 10 */
    float tan_half_fov;                                  /* tan(0.5 * 27.8 deg) in fp32, path_tracer.rs:293-302 */
} lcb_path_tracer_args;
LCB_EXPORT void lc_b200_example_path_tracer(lcb_device, lcb_stream, const lcb_path_tracer_args *, uint64_t ray_counts_out[2]);

/* Host-buffer forms: pinned staging, H2D, trace, D2H, synchronous.  These are the
 * "e2e" calls measured by bench.py. */
LCB_EXPORT void lc_b200_trace_closest_host(lcb_device, lcb_accel, const lcb_ray *rays, lcb_surface_hit *hits, uint64_t count, uint32_t mask);
LCB_EXPORT void lc_b200_trace_any_host(lcb_device, lcb_accel, const lcb_ray *rays, uint32_t *occluded, uint64_t count, uint32_t mask);

/* Instance accessors: defs::Accel vtable slots instance_transform / instance_user_id /
 * instance_visibility_mask (cpu/accel.rs:537-558, cpu/stream.rs:582-658).  Host-side,
 * synchronous with respect to completed AccelBuild commands.  `affine_out` = 12 floats row-major. */
LCB_EXPORT void lc_b200_instance_transform(lcb_device, lcb_accel, uint32_t instance, float *affine_out);
LCB_EXPORT uint32_t lc_b200_instance_user_id(lcb_device, lcb_accel, uint32_t instance);
LCB_EXPORT uint32_t lc_b200_instance_visibility_mask(lcb_device, lcb_accel, uint32_t instance);

/* Build / traversal statistics of the last build of a mesh or accel and of the last trace. */
typedef struct lcb_build_stats {
    uint64_t primitive_count;  /* triangles (mesh) or instances (accel) */
    uint64_t wide_node_count;  /* 128-byte nodes */
    uint64_t packed_tri_count; /* 64-byte leaf triangle records (36 B of vertices + prim id used) */
    uint64_t bvh_bytes;        /* nodes + packed triangles as allocated after compaction */
    uint32_t max_depth;        /* wide-tree depth */
    uint32_t was_refit;        /* 1 if PreferUpdate took the refit path */
    float build_ms;            /* device time of the last build (CUDA events) */
    uint32_t builder;          /* binary tree of the last full build: 0 LBVH split rule, 1 PLOC (see lc_b200_set_builder) */
} lcb_build_stats;
LCB_EXPORT void lc_b200_mesh_stats(lcb_device, lcb_mesh, lcb_build_stats *out);
LCB_EXPORT void lc_b200_accel_stats(lcb_device, lcb_accel, lcb_build_stats *out);

/* Instrumented traversal (same kernel compiled with counters): sums of wide nodes
 * visited and triangles tested over the batch -> the SURVEY §8(d) B_trace terms. */
typedef struct lcb_trace_counters { uint64_t nodes_visited, tris_tested, rays, instance_entries; } lcb_trace_counters;
LCB_EXPORT void lc_b200_trace_closest_counted(lcb_device, lcb_stream, lcb_accel, lcb_buffer rays, size_t rays_offset,
                                              lcb_buffer hits, size_t hits_offset, uint64_t count, uint32_t mask,
                                              lcb_trace_counters *out /* host, written after sync */);

/* Plumbing for harnesses: the cudaStream_t behind a stream, the device pointer behind a
 * buffer, the CUDA ordinal, number of kernels launched by this library so far, and a
 * version string. */
LCB_EXPORT void *lc_b200_stream_native(lcb_device, lcb_stream);
LCB_EXPORT void *lc_b200_buffer_native(lcb_device, lcb_buffer);
LCB_EXPORT int lc_b200_device_ordinal(lcb_device);
LCB_EXPORT uint64_t lc_b200_kernel_launch_count(void);
LCB_EXPORT const char *lc_b200_version(void);

/* Fabricate the `&CArc<ir::Type>` argument of create_buffer for callers that have no
 * Rust IR at hand (C/C++/Python harnesses): a struct-typed element of the given size and
 * alignment (ir.rs Type::Struct, size()/alignment() at ir.rs:325-369), or the Void type
 * (raw bytes, cpu/mod.rs:51-62) when size == 0.  Returned pointer is owned by the library
 * and lives for the process. */
LCB_EXPORT const void *lc_b200_make_ir_type(size_t size, size_t alignment);

/* Which algorithm forms the binary tree of a MeshBuild: -1 follow AccelOption.hint (FastTrace: chosen per mesh, FastBuild: LBVH),
 * 0 LBVH split rule, 1 PLOC (agglomerative clustering), 2 chosen per mesh.  Returns the previous setting.  Hits never depend on it
 * (the traversal's arithmetic is tree-independent); only build time and Mrays/s do.  Initial value: environment LC_B200_BUILDER. */
LCB_EXPORT int lc_b200_set_builder(int builder);

/* How create_shader lowers kernels that call RayTracingTraceClosest / TraceAny from their own body (csrc/ir_lower.cpp header):
 * 0 decide per kernel (wavefront lowering where it is legal and the body is short), 1 always the direct lowering (one dispatch id per
 * CUDA thread, per-thread traversal), 2 wavefront lowering wherever it is legal.  Returns the previous setting.  Results never depend on it (same arithmetic per ray); only
 * throughput does.  Initial value: environment LC_B200_LOWERING = direct | wavefront.  Applies to shaders created afterwards. */
LCB_EXPORT int lc_b200_set_lowering(int mode);

/* ---- IR -> CUDA lowering (create_shader), inspection entry points --------------------------------------------------
 * create_shader(LCKernelModule{ptr}) lowers the frontend's SSA IR (`*const ir::KernelModule`, proxy.rs:188-193;
 * layouts LC/include/luisa/rust/ir.hpp) to CUDA C++ whose ray-tracing builtins call this library's traversal
 * routines, compiles it with NVRTC for sm_100a and launches it on ShaderDispatch — the job of
 * cpu/codegen/cpp.rs + cpu/shader.rs + cpu/stream.rs:330-440 in the reference.  These three calls expose the
 * stages for tests and tooling; none of them needs a GPU.
 *   lc_b200_ir_lower_source      : the generated translation unit (malloc'd; release with LibInterface.free_string)
 *   lc_b200_shader_compile_check : lower + NVRTC compile without loading; returns 0 on success, *log (malloc'd)
 *                                  carries the lowering diagnostic or the NVRTC log
 *   lc_b200_ir_layout_json       : sizeof / offsetof / discriminants of this library's view of the IR
 *                                  (compared against the reference header's, tests/golden/ir_layout_reference.json) */
LCB_EXPORT char *lc_b200_ir_lower_source(const void *kernel_module);
LCB_EXPORT int lc_b200_shader_compile_check(const void *kernel_module, bool fast_math, char **log);
LCB_EXPORT const char *lc_b200_ir_layout_json(void);

#ifdef __cplusplus
}
#endif
#endif /* LC_B200_API_H */
