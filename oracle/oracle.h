/*
 * oracle.h — CPU restatement of the reference's ray-tracing hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under luisa-compute-rs_b200/ may link, import or call
 * this; it is used by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs as the checker and the CPU baseline.
 *
 * PARITY UNPINNED at the level of single hits: the arithmetic of the reference path lives in Embree 4
 * (crate embree_sys 0.1.11, git 6a0a591d, LC/src/rust/Cargo.lock:267-274), which is not vendored
 * under /root/reference and cannot be built here (no cargo, no clang, no Embree); the
 * reference holds no golden hit vectors for it.
 * THE REFERENCE-HELD PIN there is: luisa_compute/examples/cbox.png, the image path_tracer.rs saved — the one output of
 * MeshBuild + AccelBuild + trace_closest + trace_any + the DSL kernel around them that the reference tree contains.
 * tests/test_reference_image.py compares this file's path tracer (CPU) and the device's render through create_shader (GPU)
 * with it statistically (fixture tests/golden/cbox_reference_128.npz, generator make_cbox_reference.py): every region of the
 * image agrees within 1.5 % of its mean radiance except where one tie decides — the example's shadow rays end exactly in the
 * plane of the light quad for points of the tall box's top, and whether `t <= tmax` holds there is the last bit of t
 * (this arithmetic: 74 % occluded, an fp32 Embree-style Moeller-Trumbore emulation: 67 %, the reference image: ~23 %).
 * This file restates the *semantics* the reference wraps around Embree, file by file:
 *
 *   cpu/accel.rs:205-260   GeometryImpl::build_mesh       -> oracle_mesh_set()
 *   cpu/accel.rs:142-203   GeometryImpl::build_curve      -> oracle_curve_set()  (round curves: rounded-cone pieces cut in the frontend's
 *                          power basis, lc/src/rtx/curve.rs:88-139, LOCATE the hit — `oracle_canonical_cone()` — and a Newton iteration in
 *                          double against the true swept sphere decides it, refine_curve_hit(); hits (u, -1): accel.rs:491-494)
 *   cpu/accel.rs:324-447   AccelImpl::update              -> oracle_accel_update()
 *   cpu/accel.rs:449-509   AccelImpl::trace_closest       -> oracle_trace_closest()
 *   cpu/accel.rs:511-535   AccelImpl::trace_any           -> oracle_trace_any()
 *   cpu/accel.rs:537-558   instance_* accessors           -> oracle_instance_*()
 *   cpu/stream.rs:185-209  StreamImpl::parallel_for       -> the worker pool used by trace_*
 *   lc/src/rtx.rs:517-535  offset_ray_origin              -> oracle_offset_ray_origin()
 *
 * and fixes, where Embree leaves it open, ONE canonical per-triangle arithmetic (fp32,
 * round-to-nearest, no contraction other than the explicit fmaf calls) so that results
 * are a pure function of (scene, ray) and independent of any acceleration structure:
 * see `oracle_canonical_triangle()` in oracle.c.  The GPU kernels implement the same
 * arithmetic and must match it bit for bit on inst, prim, t and barycentrics.
 * A second, double-precision Moeller-Trumbore evaluation (`oracle_trace_closest_f64`) is
 * the geometric ground truth used to report the ambiguity ("tie") rate and to check
 * the 1e-5 relative tolerance on t / barycentrics.
 * Modes: 0 brute force (the definition), 1 binary SAH BVH (scalar, the checker for large scenes), 2 the same tree collapsed 8-wide
 * with AVX2 box tests (the CPU baseline bench.py times); all three return the same bits (tests/test_oracle.py).
 */
#ifndef LC_B200_ORACLE_H
#define LC_B200_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct __attribute__((aligned(16))) oracle_ray { float o[3], tmin, d[3], tmax; } oracle_ray; /* 32 B */
typedef struct __attribute__((aligned(8))) oracle_hit { uint32_t inst, prim; float u, v, t; uint32_t pad; } oracle_hit; /* 24 B */

typedef struct oracle_mod { /* = api::AccelBuildModification, 72 B */
    uint32_t index, user_id, flags, visibility;
    uint64_t mesh; /* oracle mesh id */
    float affine[12];
} oracle_mod;

typedef struct oracle_scene oracle_scene;

oracle_scene *oracle_scene_new(void);
void oracle_scene_free(oracle_scene *);

/* Meshes alias caller memory like Embree shared buffers (accel.rs:217-237): the oracle keeps
 * the pointers and re-reads them at every commit.  Returns mesh id. */
uint64_t oracle_mesh_new(oracle_scene *);
void oracle_mesh_set(oracle_scene *, uint64_t mesh, const void *vertices, size_t vertex_stride, size_t vertex_count,
                     const void *indices, size_t index_stride, size_t triangle_count);
/* (Re)build the per-mesh CPU BVH after the vertex data changed. */
void oracle_mesh_commit(oracle_scene *, uint64_t mesh);
/* cpu/accel.rs:142-203 GeometryImpl::build_curve: a curve geometry on a handle from oracle_mesh_new(); basis = api::CurveBasis
 * (0 PiecewiseLinear, 1 CubicBSpline, 2 CatmullRom, 3 Bezier); control points float4 {x, y, z, radius} at cp_stride; buffers are aliased. */
void oracle_curve_set(oracle_scene *, uint64_t mesh, int basis, const void *cps, size_t cp_stride, size_t cp_count, const uint32_t *segs, size_t seg_count);
int oracle_canonical_cone(const float o[3], const float d[3], float tmin, float tmax, const float A[4], const float B[4], float *t, float *s);

/* AccelImpl::update (accel.rs:324-447). */
void oracle_accel_update(oracle_scene *, uint32_t instance_count, const oracle_mod *mods, size_t n_mods);

void oracle_instance_transform(const oracle_scene *, uint32_t inst, float affine[12]);
uint32_t oracle_instance_user_id(const oracle_scene *, uint32_t inst);
uint32_t oracle_instance_visibility(const oracle_scene *, uint32_t inst);
uint32_t oracle_instance_count(const oracle_scene *);

/* mode: 0 = brute force over every triangle of every instance (definition),
 *       1 = CPU BVH (same results, checked against mode 0 in tests).
 * threads: worker count (<=0: all online cores); workers pull 64-ray blocks from an atomic
 * counter, the scheme of StreamImpl::parallel_for. */
void oracle_trace_closest(const oracle_scene *, const oracle_ray *rays, uint64_t n, uint32_t mask, oracle_hit *hits, int mode, int threads);
void oracle_trace_any(const oracle_scene *, const oracle_ray *rays, uint64_t n, uint32_t mask, uint32_t *occluded, int mode, int threads);

/* AccelImpl::ray_query (accel.rs:582-800) in batch form, triangles only: opaque instances commit their hits without
 * a callback, every triangle candidate of a NON-opaque instance (accel.rs:378-394 enables the filter only there) is
 * shown to the candidate hook with its canonical fp32 barycentrics and counts only if the hook commits.  The hook is
 * the same small set of pure predicates the device offers (include/lc_b200_api.h LCB_FILTER_*).  NB the reference CPU
 * backend records a committed hit only inside its callbacks, so there opaque triangles never reach `rq.hit`; this
 * restatement follows the frontend's documented semantics (and the DX/OptiX backends): opaque triangles commit.
 * terminate_on_first = RayTracingQueryAny (rtcOccluded1): WHICH hit is reported is then traversal-order dependent.
 * Nothing committed -> {~0, ~0, (0,0), hit_type 0, t 0} (the zero-initialised RayQuery of cpu_resource.h:320-331). */
typedef struct __attribute__((aligned(8))) oracle_committed_hit { uint32_t inst, prim; float u, v; uint32_t hit_type; float t; } oracle_committed_hit;
/* kind 0 commit all, 1 barycentric disc (examples/ray_query.rs:148-162), 2 per-primitive bit table, 3 reject all,
 * 4 stripes (examples/path_tracer_cutout.rs:364-373): on instance i with stripe_freq[i] != 0 a candidate commits iff
 * fract(bary.y * stripe_freq[i]) < stripe_keep[i], fract(x) = x - floorf(x) (device_math.h lc_fract). */
typedef struct oracle_filter { int kind; float radius; const uint32_t *bits; const uint32_t *first_bit; const float *stripe_freq; const float *stripe_keep; } oracle_filter;
void oracle_ray_query(const oracle_scene *, const oracle_ray *rays, uint64_t n, uint32_t mask, int terminate_on_first, const oracle_filter *filter,
                      oracle_committed_hit *out, int mode, int threads);

/* CPU restatement of the kernel of luisa_compute/examples/path_tracer.rs:247-455 over this oracle's trace functions (one
 * dispatch: spp_per_dispatch samples per pixel, image += radiance, w += 1, seeds advanced).  Same operation order as
 * luisa-compute-rs_b200/csrc/path_tracer.cu (which is compiled without FMA contraction), same fixed sin/cos polynomial,
 * so the two images agree bit for bit.  vertex_heap[i] / index_heap[i]: tightly packed [f32;3] / [u32;3] of instance i. */
void oracle_path_tracer(const oracle_scene *, const float *const *vertex_heap, const uint32_t *const *index_heap, float *image_rgba, uint32_t *seed_image,
                        uint32_t width, uint32_t height, uint32_t spp_per_dispatch, uint32_t max_depth, float tan_half_fov, int threads,
                        uint64_t ray_counts_out[2]);

/* examples/path_tracer_cutout.rs: the same kernel with `accel.traverse(ray).on_surface_hit(|c| if filter(&c) { c.commit() }).trace()` in
 * place of intersect and `traverse_any` in place of intersect_any (:379-389, :435-445); every instance is pushed non-opaque (:250), so every
 * triangle candidate passes through `filter` (an oracle_filter, kind 4 for the example's stripes on the two boxes). */
void oracle_path_tracer_cutout(const oracle_scene *, const float *const *vertex_heap, const uint32_t *const *index_heap, float *image_rgba, uint32_t *seed_image,
                               uint32_t width, uint32_t height, uint32_t spp_per_dispatch, uint32_t max_depth, float tan_half_fov, const oracle_filter *filter,
                               int threads, uint64_t ray_counts_out[2]);

/* Double-precision ground truth.  ambiguous[i] (may be NULL) is set to 1 when the fp32
 * answer is not forced: a second candidate lies within rel 1e-6 of the closest t, or some
 * triangle's nearest barycentric (hit or near miss, |b| < 1e-5) is within rounding of an edge
 * at a t that could change the answer. */
void oracle_trace_closest_f64(const oracle_scene *, const oracle_ray *rays, uint64_t n, uint32_t mask, oracle_hit *hits, uint8_t *ambiguous, int mode, int threads);

/* lc/src/rtx.rs:517-535 */
void oracle_offset_ray_origin(const float p[3], const float n[3], float out[3]);

/* The canonical single ray/triangle evaluation, exported for unit tests.
 * Returns 1 and fills t,u,v when tmin < t <= tmax. */
int oracle_canonical_triangle(const float o[3], const float d[3], float tmin, float tmax,
                              const float v0[3], const float v1[3], const float v2[3], float *t, float *u, float *v);

/* world->object 3x4 (row-major) from the instance affine: double adjugate inverse rounded to fp32. */
void oracle_invert_affine(const float m[12], float inv[12]);

int oracle_hw_threads(void);
#ifdef __cplusplus
}
#endif
#endif
