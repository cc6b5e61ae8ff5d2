// device.cu — the far side of luisa-compute-rs's DeviceInterface for the "b200" device.
//
// Mirrors the behaviour of the reference CPU backend object model
// (luisa_compute_backend_impl/src/cpu/{mod.rs,stream.rs,accel.rs,resource.rs}) for the hot path:
// handles are raw pointers to backend objects, buffers are zero-initialised device memory,
// streams execute command lists in order and fire the completion callback exactly once from a
// stream-owned host thread, MeshBuild / AccelBuild run the CUDA builder, and every failure is
// logged through the logger callback and aborts (backend_impl/src/lib.rs:101-131).
// Slots outside the hot path log "unsupported" and abort — there is no CPU fallback anywhere.
#include "../../include/lc_b200_api.h"
#include "build.cuh"
#include "shader.h"

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <dlfcn.h>
#include <nvtx3/nvToolsExt.h>
#include <chrono>
#include <map>
#include <mutex>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

using namespace lcb;

namespace {

// NVTX ranges around build stages and dispatches (what the reference's CUDA backend marks in cuda_primitive.cpp:22,53,84,113): visible to
// nsys / ncu --nvtx, free when no tool is attached (nvtx3 is header-only and resolves the injection library lazily).
struct nvtx_range {
    explicit nvtx_range(const char *name) { nvtxRangePushA(name); }
    ~nvtx_range() { nvtxRangePop(); }
};

// ---- logging / errors ------------------------------------------------------------------------
std::atomic<void (*)(lcb_logger_message)> g_logger{nullptr};
std::atomic<unsigned long long> g_launches{0};
// A device keeps more than eight streams alive (the user's, its own copy / build / event lanes).  With CUDA's default of eight hardware
// queues they alias, and a stream parked on a timeline value that is signalled later (wait before signal: legal, cpu/resource.rs:10-44) can
// then hold up the stream that is to signal it — measured as a hang of the e2e host program.  Effective only when this
// library is loaded before the process creates its CUDA context, which is the case for a luisa-compute-rs program.
const int g_more_queues = setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0);
// -1: follow AccelOption.hint; otherwise kBuilderLbvh / kBuilderPloc / kBuilderAuto for every mesh (LC_B200_BUILDER, lc_b200_set_builder)
std::atomic<int> g_builder_override{[] { const char *e = getenv("LC_B200_BUILDER"); return !e ? -1 : (strcmp(e, "ploc") == 0 ? 1 : (strcmp(e, "auto") == 0 ? 2 : (strcmp(e, "lbvh") == 0 ? 0 : -1))); }()};

void log_msg(const char *level, const char *fmt, ...) {
    char buf[2048];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof(buf), fmt, ap); va_end(ap);
    auto cb = g_logger.load();
    if (cb) { lcb_logger_message m{"lc_b200", level, buf}; cb(m); }
    else if (level[0] == 'E' || level[0] == 'W' || getenv("LC_B200_LOG")) fprintf(stderr, "[lc_b200][%s] %s\n", level, buf);
}

[[noreturn]] void fatal(const char *fmt, ...) {
    char buf[2048];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof(buf), fmt, ap); va_end(ap);
    log_msg("E", "%s", buf);
    fprintf(stderr, "[lc_b200] fatal: %s\n", buf);
    fflush(stderr);
    abort();
}

#define CUDA_CHECK(expr)                                                                                   \
    do {                                                                                                   \
        cudaError_t err__ = (expr);                                                                        \
        if (err__ != cudaSuccess) fatal("CUDA error %s at %s:%d: %s", cudaGetErrorName(err__), __FILE__, __LINE__, cudaGetErrorString(err__)); \
    } while (0)

// ---- IR type walking (create_buffer's &CArc<ir::Type>) ----------------------------------------
// Layouts: LC/include/luisa/rust/ir_common.h (CArcSharedBlock, CBoxedSlice) and ir.hpp:187-232.
struct IrArcBlock { void *ptr; std::atomic<size_t> ref_count; void (*destructor)(void *); };
struct IrSlice { void *ptr; size_t len; void *destructor; };
enum IrTypeTag : int32_t { IR_VOID, IR_USERDATA, IR_PRIMITIVE, IR_VECTOR, IR_MATRIX, IR_STRUCT, IR_ARRAY, IR_OPAQUE };
struct IrVectorElement { int32_t tag; int32_t pad; union { int32_t primitive; IrArcBlock *vector; } u; };
struct IrVectorType { IrVectorElement element; uint32_t length; };
struct IrType {
    int32_t tag; int32_t pad;
    union {
        int32_t primitive;
        IrVectorType vector;  // also MatrixType{element, dimension}
        struct { IrSlice fields; size_t alignment; size_t size; } struct_;
        struct { IrArcBlock *element; size_t length; } array;
    } u;
};

size_t ir_primitive_size(int32_t p) {  // ir.rs:214-231
    static const size_t sz[12] = {1, 1, 1, 2, 2, 4, 4, 8, 8, 2, 4, 8};
    if (p < 0 || p >= 12) fatal("bad IR primitive %d", p);
    return sz[p];
}
size_t ir_vector_size(const IrVectorType &v);
size_t ir_vector_element_size(const IrVectorElement &e) {
    return e.tag == 0 ? ir_primitive_size(e.u.primitive) : ir_vector_size(*(const IrVectorType *)e.u.vector->ptr);
}
size_t ir_vector_size(const IrVectorType &v) {  // ir.rs:234-251: 3-vectors of scalars are padded to 4
    size_t el = ir_vector_element_size(v.element);
    uint32_t len = v.length;
    if (v.element.tag == 0) { uint32_t four = len / 4, rem = len % 4; len = rem <= 2 ? four * 4 + rem : four * 4 + 4; }
    return el * len;
}
size_t ir_type_size(const IrType *t);
size_t ir_type_alignment(const IrType *t) {  // ir.rs:356-369
    switch (t->tag) {
        case IR_VOID: case IR_USERDATA: return 0;
        case IR_PRIMITIVE: return ir_primitive_size(t->u.primitive);
        case IR_STRUCT: return t->u.struct_.alignment;
        case IR_VECTOR: case IR_MATRIX: {
            size_t el = ir_primitive_size(t->u.vector.element.u.primitive);
            uint32_t dim = t->u.vector.length == 3 ? 4 : t->u.vector.length;
            size_t a = el * dim; return a < 16 ? a : 16;
        }
        case IR_ARRAY: return ir_type_alignment((const IrType *)t->u.array.element->ptr);
        default: fatal("unsupported IR type tag %d", t->tag);
    }
}
size_t ir_type_size(const IrType *t) {  // ir.rs:325-335
    switch (t->tag) {
        case IR_VOID: case IR_USERDATA: return 0;
        case IR_PRIMITIVE: return ir_primitive_size(t->u.primitive);
        case IR_STRUCT: return t->u.struct_.size;
        case IR_VECTOR: return ir_vector_size(t->u.vector);
        case IR_MATRIX: {  // ir.rs:275-293
            uint32_t d = t->u.vector.length;
            uint32_t cols = d == 2 ? 2u : 4u;
            return ir_primitive_size(t->u.vector.element.u.primitive) * cols * d;
        }
        case IR_ARRAY: return ir_type_size((const IrType *)t->u.array.element->ptr) * t->u.array.length;
        default: fatal("unsupported IR type tag %d", t->tag);
    }
}

// Zero-fill of a fresh allocation.  cudaMemset runs on the legacy default stream and returns before it has executed; the
// device's streams are non-blocking (they do not order against that stream), so an upload submitted right after creation
// could be overtaken by the fill.  Waiting for the fill here closes that window.
static void zero_fill(void *p, size_t bytes) {
    CUDA_CHECK(cudaMemsetAsync(p, 0, bytes, cudaStreamLegacy));
    CUDA_CHECK(cudaStreamSynchronize(cudaStreamLegacy));
}

// ---- backend objects ---------------------------------------------------------------------------
struct BufferObj { uint8_t *ptr = nullptr; size_t size = 0; bool owned = true; };

// A build command is enqueued and dispatch() returns (cpu/mod.rs:168-180 enqueues and returns; cuda_primitive.cpp:20-110 builds on the
// stream): what the host wants to know about the finished build — node count, depth, the builder's error flag, the device time — is
// copied into pinned memory behind the build and picked up lazily, when somebody asks (stats, the next build of the same object, a
// refit that needs the node count, destroy).
struct PendingBuild {
    bool active = false;
    BuildHeader *h_hdr = nullptr;   // pinned
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    void begin(cudaStream_t st) {
        if (!h_hdr) {
            CUDA_CHECK(cudaHostAlloc((void **)&h_hdr, sizeof(BuildHeader), cudaHostAllocDefault));
            CUDA_CHECK(cudaEventCreate(&e0)); CUDA_CHECK(cudaEventCreate(&e1));
        }
        memset(h_hdr, 0, sizeof(BuildHeader));
        CUDA_CHECK(cudaEventRecord(e0, st));
    }
    void end(cudaStream_t st, const BuildHeader *device_header) {
        if (device_header) CUDA_CHECK(cudaMemcpyAsync(h_hdr, device_header, sizeof(BuildHeader), cudaMemcpyDeviceToHost, st));
        CUDA_CHECK(cudaEventRecord(e1, st));
        active = true;
    }
    float wait_ms() {  // blocks until the build has finished on the device
        CUDA_CHECK(cudaEventSynchronize(e1));
        float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
        active = false;
        return ms;
    }
    void destroy() {
        if (h_hdr) { cudaFreeHost(h_hdr); cudaEventDestroy(e0); cudaEventDestroy(e1); h_hdr = nullptr; }
    }
};

struct MeshObj {
    lcb_accel_option option{};
    bool built = false;
    int auto_builder = -1; uint32_t auto_builder_n = 0;  // what the per-mesh builder choice picked last time, and for how many triangles
    bool procedural = false;  // created by create_procedural_primitive: leaves are user AABBs, hits come from the RayQuery callback
    bool curve = false;       // created by create_curve: leaves are the rounded-cone pieces of the segments (trace_device.cuh "curves")
    uint32_t n_tris = 0;
    uint64_t generation = 0;  // bumped whenever nodes/tris are re-allocated
    WideNode *nodes = nullptr; PackedTri *tris = nullptr;
    uint32_t n_nodes = 0, n_packed = 0, node_capacity = 0;
    RefitArrays refit{nullptr, nullptr, nullptr};  // side arrays of the refit path, allocated at the first PreferUpdate
    lcb_build_stats stats{};
    PendingBuild pend;                             // the last MeshBuild as the stream will finish it (mesh_finalize folds it in)
    bool pend_refit = false; int pend_builder = 0;
    std::mutex mu;
};

struct InstanceHost {  // AccelImpl::Instance, cpu/accel.rs:270-296
    float affine[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
    uint32_t user_id = 0, visible = 0xff;
    bool opaque = true, valid = false;
    MeshObj *mesh = nullptr;
    uint64_t mesh_generation = 0;
};

struct AccelObj {
    lcb_accel_option option{};
    std::vector<InstanceHost> instances;
    InstanceRec *table = nullptr; uint32_t table_capacity = 0;
    WideNode *tlas_nodes = nullptr; uint32_t *tlas_prims = nullptr; uint32_t *active_ids = nullptr; uint32_t tlas_capacity = 0;
    uint32_t n_active = 0;
    float world_lo[3] = {0, 0, 0}, world_hi[3] = {0, 0, 0};
    uint32_t *dirty = nullptr;  // device flag raised by kernels that edit the instance table (lc_set_instance_*)
    uint8_t *h_stage = nullptr; size_t h_stage_cap = 0;  // pinned staging of modification records + active ids (grow-only; reused once the previous build has finished)
    lcb_build_stats stats{};
    PendingBuild pend;          // the last AccelBuild as the stream will finish it (accel_finalize folds it in)
    uint32_t pend_instances = 0, pend_active = 0; bool pend_table_only = false;
    bool maybe_dirty = false;   // a kernel that can edit the instance table (RayTracingSetInstance*) was dispatched with this accel since the last build
    std::mutex mu;
};

// Textures: one mip level, texels row-major in device memory (the CPU backend tiles its textures, cpu/texture.rs; uploads and
// downloads present the same row-major image to the user either way).  storage = api PixelStorage (api_types:366-383).
struct TextureObj { uint8_t *ptr = nullptr; uint32_t dim = 2, width = 1, height = 1, depth = 1; int32_t storage = 0; size_t pixel_bytes = 4, bytes = 0; };

// Bindless arrays (cpu/resource.rs:60-124): a slot table mirrored on the host; BindlessArrayUpdate uploads touched slots.
struct BindlessObj { std::vector<HostBindlessSlot> host; HostBindlessSlot *device = nullptr; std::mutex mu; };

// Timeline event (cpu/resource.rs:10-44): one monotone 64-bit counter in device memory.  signal = a one-thread kernel doing atomicMax
// on the signalling stream (EventImpl::signal is fetch_max), wait = cuStreamWaitValue64(>=) enqueued on the waiting stream — the host
// never blocks, so wait-before-signal works as on the reference's stream threads (cpu/mod.rs:367-402) — synchronize / is_completed
// read the counter back through a private copy stream.  `seen` caches the largest value the host has observed.
struct EventObj {
    unsigned long long *counter = nullptr;
    std::atomic<uint64_t> seen{0};
};

struct DeviceObj;

// Upload snapshots.  BufferUpload / TextureUpload sources are only borrowed until dispatch() returns (the frontend's copy_from_async
// borrows for the life of the Command; the CPU backend memcpy's them into staging buffers inside dispatch, cpu/stream.rs:33-64), so the
// bytes are copied into pinned staging blocks before dispatch returns — whatever the source is (pageable, pinned, registered: a pinned
// source would otherwise be read by the copy engine long after the caller may have reused it) — and the H2D copy runs from the staging
// block.  Blocks are pooled by size class (cudaHostAlloc costs milliseconds) and large snapshots are taken by several host threads.
struct PinnedPool {
    std::mutex mu;
    std::multimap<size_t, void *> free_blocks;
    size_t pooled_bytes = 0;
    static size_t size_class(size_t n) { size_t c = 64 << 10; while (c < n) c <<= 1; return c; }
    void *take(size_t n, size_t &cap) {
        cap = size_class(n);
        {
            std::lock_guard<std::mutex> lk(mu);
            auto it = free_blocks.find(cap);
            if (it != free_blocks.end()) { void *p = it->second; free_blocks.erase(it); pooled_bytes -= cap; return p; }
        }
        void *p = nullptr;
        CUDA_CHECK(cudaHostAlloc(&p, cap, cudaHostAllocDefault));
        return p;
    }
    void give(void *p, size_t cap) {
        {
            std::lock_guard<std::mutex> lk(mu);
            if (pooled_bytes + cap <= (size_t(4) << 30)) { free_blocks.emplace(cap, p); pooled_bytes += cap; return; }
        }
        cudaFreeHost(p);
    }
    void clear() {
        std::lock_guard<std::mutex> lk(mu);
        for (auto &kv : free_blocks) cudaFreeHost(kv.second);
        free_blocks.clear(); pooled_bytes = 0;
    }
};
PinnedPool g_pinned;
void release_staged(std::vector<std::pair<void *, size_t>> &v) { for (auto &b : v) g_pinned.give(b.first, b.second); v.clear(); }

void parallel_copy(void *dst, const void *src, size_t n) {
    static const unsigned hw = std::max(1u, std::min(8u, std::thread::hardware_concurrency()));
    if (n < (size_t(8) << 20) || hw == 1) { memcpy(dst, src, n); return; }
    const size_t slice = ((n + hw - 1) / hw + 4095) & ~size_t(4095);
    std::vector<std::thread> workers;
    for (size_t off = slice; off < n; off += slice) workers.emplace_back([=] { memcpy((uint8_t *)dst + off, (const uint8_t *)src + off, std::min(slice, n - off)); });
    memcpy(dst, src, std::min(slice, n));
    for (auto &w : workers) w.join();
}

struct StreamObj {
    DeviceObj *dev = nullptr;
    cudaStream_t stream = nullptr;
    unsigned long long *work_counter = nullptr;   // device: ray-pool counter for trace launches on this stream
    struct CopyOut { void *dst; void *stage; size_t bytes, cap; };  // a download into pageable memory: pinned staging -> destination, by the stream's worker
    struct Pending { cudaEvent_t ev; lcb_dispatch_callback cb; uint8_t *ctx; std::vector<std::pair<void *, size_t>> staged; /* pinned upload snapshots, back to the pool on completion */
                     std::vector<CopyOut> copy_out; };
    std::mutex mu; std::condition_variable cv, drained;
    std::deque<Pending> pending;
    bool stop = false; size_t in_flight = 0;
    std::thread worker;

    void run() {
        for (;;) {
            Pending p;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return stop || !pending.empty(); });
                if (pending.empty()) return;
                p = std::move(pending.front()); pending.pop_front();
            }
            cudaError_t e = cudaEventSynchronize(p.ev);
            if (e != cudaSuccess) fatal("stream failed: %s", cudaGetErrorString(e));
            cudaEventDestroy(p.ev);
            for (auto &c : p.copy_out) { parallel_copy(c.dst, c.stage, c.bytes); g_pinned.give(c.stage, c.cap); }
            release_staged(p.staged);
            if (p.cb) p.cb(p.ctx);
            { std::lock_guard<std::mutex> lk(mu); in_flight--; }
            drained.notify_all();
        }
    }
    void push(Pending &&p) {
        { std::lock_guard<std::mutex> lk(mu); pending.push_back(std::move(p)); in_flight++; }
        cv.notify_all();
    }
    void wait_drained() { std::unique_lock<std::mutex> lk(mu); drained.wait(lk, [&] { return in_flight == 0; }); }
};

struct DeviceObj {
    int ordinal = 0;
    LaunchCounter lc;
    StreamObj *internal = nullptr, *internal2 = nullptr;  // traversal lanes of the *_host entry points (alternating chunks)
    cudaStream_t copy_in = nullptr, copy_out = nullptr;  // H2D / D2H lanes of the pipelined host entry points
    // grow-only device staging for the host entry points
    uint8_t *stage_rays = nullptr, *stage_out = nullptr; size_t stage_rays_cap = 0, stage_out_cap = 0;
    std::mutex mu;
    // Grow-only arena for BLAS builds: the builder's scratch and the full-capacity node array a build is collapsed into before
    // compaction.  A MeshBuild ends synchronised, so one arena serves every mesh of the device (build_mu serialises them);
    // rebuilding the same scene every frame (C4) then touches no allocator at all.  Stream-ordered pool allocations of these
    // multi-GB blocks made a rebuild after a few refits cost 15-30 ms instead of 8 (tools/micro/rebuild_probe.py).
    uint8_t *build_arena = nullptr; size_t build_arena_cap = 0;
    cudaEvent_t arena_event = nullptr;   // recorded behind the last build that used the arena: the next build (any stream) waits for it on the device
    std::mutex build_mu;
    // event counters are read back on their own stream (never behind the user's work)
    cudaStream_t poll_stream = nullptr; unsigned long long *poll_word = nullptr; std::mutex poll_mu;
};

template <class T> T *as(uint64_t h) { if (h == 0 || h == LCB_INVALID_HANDLE) fatal("invalid resource handle"); return reinterpret_cast<T *>(h); }
DeviceObj *dev_of(lcb_device d) { return as<DeviceObj>(d.id); }
void bind(DeviceObj *d) { CUDA_CHECK(cudaSetDevice(d->ordinal)); }
void flush_launches(DeviceObj *d) { g_launches += d->lc.count.exchange(0); }

// ---- buffers -----------------------------------------------------------------------------------
lcb_created_buffer create_buffer(lcb_device dev, const void *ir_type, size_t count, void *ext_mem) {
    DeviceObj *d = dev_of(dev); bind(d);
    if (!ir_type) fatal("create_buffer: null type");
    const IrArcBlock *blk = *reinterpret_cast<IrArcBlock *const *>(ir_type);
    const IrType *ty = blk ? reinterpret_cast<const IrType *>(blk->ptr) : nullptr;
    if (!ty) fatal("create_buffer: null CArc<Type>");
    const size_t stride = ir_type_size(ty);
    const size_t total = ty->tag == IR_VOID ? count : stride * count;  // cpu/mod.rs:57-61
    auto *b = new BufferObj;
    b->size = total;
    if (ext_mem) { b->ptr = (uint8_t *)ext_mem; b->owned = false; }
    else {
        CUDA_CHECK(cudaMalloc(&b->ptr, total ? total : 16));
        zero_fill(b->ptr, total ? total : 16);  // BufferImpl::new zero-initialises (cpu/resource.rs:126-133)
    }
    lcb_created_buffer out{};
    out.resource.handle = (uint64_t)b; out.resource.native_handle = b->ptr;
    out.element_stride = stride; out.total_size_bytes = total;
    return out;
}
void destroy_buffer(lcb_device dev, lcb_buffer h) {
    DeviceObj *d = dev_of(dev); bind(d);
    BufferObj *b = as<BufferObj>(h.id);
    if (b->owned) CUDA_CHECK(cudaFree(b->ptr));
    delete b;
}

// ---- streams -----------------------------------------------------------------------------------
StreamObj *make_stream(DeviceObj *d) {
    auto *s = new StreamObj; s->dev = d;
    CUDA_CHECK(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
    CUDA_CHECK(cudaMalloc(&s->work_counter, 256));
    s->worker = std::thread([s] { s->run(); });
    return s;
}
lcb_created create_stream(lcb_device dev, int32_t) {
    DeviceObj *d = dev_of(dev); bind(d);
    StreamObj *s = make_stream(d);
    return lcb_created{(uint64_t)s, (void *)s->stream};
}
void synchronize_stream(lcb_device dev, lcb_stream h) {
    DeviceObj *d = dev_of(dev); bind(d);
    StreamObj *s = as<StreamObj>(h.id);
    CUDA_CHECK(cudaStreamSynchronize(s->stream));
    s->wait_drained();
}
void free_stream(StreamObj *s) {
    CUDA_CHECK(cudaStreamSynchronize(s->stream));
    s->wait_drained();
    { std::lock_guard<std::mutex> lk(s->mu); s->stop = true; }
    s->cv.notify_all();
    s->worker.join();
    cudaFree(s->work_counter);
    cudaStreamDestroy(s->stream);
    delete s;
}
void destroy_stream(lcb_device dev, lcb_stream h) { DeviceObj *d = dev_of(dev); bind(d); free_stream(as<StreamObj>(h.id)); }

// ---- mesh build (GeometryImpl::build_mesh, cpu/accel.rs:205-260) ---------------------------------
void blas_build(DeviceObj *d, StreamObj *s, MeshObj *m, uint32_t n, int32_t request, const TriangleInput &in, const uint8_t *aabbs, const CurveInput *curve = nullptr);
void accel_finalize(AccelObj *a);

void mesh_build(DeviceObj *d, StreamObj *s, const lcb_cmd_mesh_build &c) {
    MeshObj *m = as<MeshObj>(c.mesh.id);
    if (m->procedural || m->curve) fatal("MeshBuild on a %s handle", m->curve ? "curve" : "procedural primitive");
    std::lock_guard<std::mutex> lk(m->mu);
    if (c.index_stride != 12) fatal("Index stride must be 12 (got %zu).", c.index_stride);  // api/runtime.cpp:191-193
    if (c.vertex_stride < 12) fatal("vertex stride must be >= 12 (got %zu)", c.vertex_stride);
    BufferObj *vb = as<BufferObj>(c.vertex_buffer.id), *ib = as<BufferObj>(c.index_buffer.id);
    if (c.vertex_buffer_offset + c.vertex_buffer_size > vb->size) fatal("MeshBuild: vertex range exceeds buffer");
    if (c.index_buffer_offset + c.index_buffer_size > ib->size) fatal("MeshBuild: index range exceeds buffer");
    const uint32_t n = (uint32_t)(c.index_buffer_size / c.index_stride);
    TriangleInput in{vb->ptr + c.vertex_buffer_offset, c.vertex_stride, ib->ptr + c.index_buffer_offset};
    blas_build(d, s, m, n, c.request, in, nullptr);
}

// ProceduralPrimitiveBuild (GeometryImpl::build_procedural, cpu/accel.rs:84-141): a BLAS over the user's AABBs; always a full build.
void procedural_build(DeviceObj *d, StreamObj *s, const lcb_cmd_procedural_build &c) {
    MeshObj *m = as<MeshObj>(c.handle.id);
    if (!m->procedural) fatal("ProceduralPrimitiveBuild on a mesh handle");
    std::lock_guard<std::mutex> lk(m->mu);
    BufferObj *ab = as<BufferObj>(c.aabb_buffer.id);
    if (c.aabb_offset + c.aabb_count * 24 > ab->size) fatal("ProceduralPrimitiveBuild: AABB range exceeds buffer");
    blas_build(d, s, m, (uint32_t)c.aabb_count, LCB_REQUEST_FORCE_BUILD, TriangleInput{nullptr, 0, nullptr}, ab->ptr + c.aabb_offset);
}

// CurveBuild (GeometryImpl::build_curve, cpu/accel.rs:142-203): control points are read as float4 {x, y, z, radius} at cp_stride
// (>= 16, asserted there), segments are u32 first-control-point indices.  The reference refits on PreferUpdate (rtcUpdateGeometryBuffer);
// a curve BLAS is small next to a mesh, so every CurveBuild here is a full build — same result.
void curve_build(DeviceObj *d, StreamObj *s, const lcb_cmd_curve_build &c) {
    MeshObj *m = as<MeshObj>(c.curve.id);
    if (!m->curve) fatal("CurveBuild on a handle that is not a curve");
    std::lock_guard<std::mutex> lk(m->mu);
    if (c.basis < 0 || c.basis > 3) fatal("CurveBuild: unknown basis %d", c.basis);
    if (c.cp_stride < 16) fatal("cp buffer stride must be >= 16 (got %zu)", c.cp_stride);  // cpu/accel.rs:159
    if (c.cp_stride % 16 != 0) fatal("CurveBuild: cp buffer stride must be a multiple of 16 (got %zu)", c.cp_stride);
    BufferObj *cb = as<BufferObj>(c.cp_buffer.id), *sb = as<BufferObj>(c.seg_buffer.id);
    if (c.cp_offset % 16 != 0) fatal("CurveBuild: cp buffer offset must be a multiple of 16");
    if (c.cp_count && c.cp_offset + (c.cp_count - 1) * c.cp_stride + 16 > cb->size) fatal("CurveBuild: control point range exceeds buffer");
    if (c.seg_offset % 4 != 0 || c.seg_offset + c.seg_count * 4 > sb->size) fatal("CurveBuild: segment range exceeds buffer");
    const uint32_t pieces = c.basis == 0 ? 1u : kCurveSubdiv, per_seg = c.basis == 0 ? 2u : 4u;
    if (c.seg_count * pieces > 0x7fffffffull) fatal("CurveBuild: too many segments");
    if (c.seg_count) {  // every segment must stay inside the control points (Embree reads them unchecked)
        std::vector<uint32_t> segs(c.seg_count);
        CUDA_CHECK(cudaMemcpyAsync(segs.data(), sb->ptr + c.seg_offset, c.seg_count * 4, cudaMemcpyDeviceToHost, s->stream));
        CUDA_CHECK(cudaStreamSynchronize(s->stream));
        for (size_t i = 0; i < c.seg_count; i++)
            if ((size_t)segs[i] + per_seg > c.cp_count) fatal("CurveBuild: segment %zu starts at control point %u of %zu", i, segs[i], c.cp_count);
    }
    CurveInput in{cb->ptr + c.cp_offset, c.cp_stride, reinterpret_cast<const uint32_t *>(sb->ptr + c.seg_offset), (uint32_t)c.basis, pieces};
    blas_build(d, s, m, (uint32_t)(c.seg_count * pieces), LCB_REQUEST_FORCE_BUILD, TriangleInput{nullptr, 0, nullptr}, nullptr, &in);
}

// Fold a finished MeshBuild into the host-side record (PendingBuild comment).  Caller holds m->mu.
void mesh_finalize(MeshObj *m) {
    if (!m->pend.active) return;
    const float ms = m->pend.wait_ms();
    m->stats.build_ms = ms;
    if (m->pend_refit) { m->stats.was_refit = 1; return; }
    const BuildHeader &hdr = *m->pend.h_hdr;
    const uint32_t n = m->n_tris;
    if (hdr.error) fatal("BVH build failed (code %u: %s)", hdr.error, hdr.error == 1 ? "tree deeper than the traversal stack" : "node capacity exceeded");
    if (hdr.emitted != n || hdr.prim_count != n) fatal("BVH build inconsistent: emitted %u of %u primitives", hdr.emitted, n);
    m->n_nodes = hdr.node_count; m->n_packed = hdr.prim_count;
    m->stats.wide_node_count = hdr.node_count; m->stats.packed_tri_count = hdr.prim_count;
    m->stats.bvh_bytes = (uint64_t)hdr.node_count * sizeof(WideNode) + (uint64_t)n * sizeof(PackedTri);
    m->stats.max_depth = hdr.max_depth; m->stats.was_refit = 0; m->stats.builder = (uint32_t)m->pend_builder;
}

// Wide nodes a tree over n primitives can have.  Every wide node absorbs at least min(7, kLeafMax) = 2 internal nodes of the binary
// tree (k_collapse opens children until eight, or until every child is a single primitive; a subtree of <= kLeafMax primitives is a
// leaf child, not a node), and the binary tree has n - 1 of them: at most (n - 1) / 2 wide nodes, the root included.  The node block is
// allocated for that bound up front, which is what lets the build run without the host (no count to wait for, no compaction copy).
static uint32_t wide_node_bound(uint32_t n) { return n / 2 + 2; }

void blas_build(DeviceObj *d, StreamObj *s, MeshObj *m, uint32_t n, int32_t request, const TriangleInput &in, const uint8_t *aabbs, const CurveInput *curve) {
    cudaStream_t st = s->stream;
    mesh_finalize(m);  // the previous build of this mesh (usually long finished): its node count is needed below, its record is reused
    // PreferUpdate on a built, updatable mesh is a vertex-update refit (accel.rs:251-257; the OptiX backend's rule —
    // rebuild when updates are not allowed, the mesh was never built or its size changed — is cuda_mesh.cpp:42-68).
    // The BVH aliases the user buffers like Embree's shared geometry buffers (accel.rs:217-237): vertices are re-read now.
    if (request == LCB_REQUEST_PREFER_UPDATE && m->built && m->option.allow_update && n == m->n_tris && n > 0 && m->nodes) {
        nvtx_range range("lc_b200 MeshBuild (refit)");
        m->pend.begin(st);
        if (!m->refit.parent) {
            CUDA_CHECK(cudaMallocAsync((void **)&m->refit.parent, (size_t)m->n_nodes * 4, st));
            CUDA_CHECK(cudaMallocAsync((void **)&m->refit.boxes, (size_t)m->n_nodes * 24, st));
            CUDA_CHECK(cudaMallocAsync((void **)&m->refit.counters, (size_t)m->n_nodes * 4, st));
            build_refit_arrays(st, m->n_nodes, m->nodes, m->refit, d->lc);
        }
        refit_blas(st, m->n_nodes, n, in, m->nodes, m->tris, m->refit, nullptr, d->lc);
        m->pend_refit = true;
        m->pend.end(st, nullptr);
        return;
    }
    nvtx_range range("lc_b200 MeshBuild (build)");
    if (m->refit.parent) {
        CUDA_CHECK(cudaFreeAsync(m->refit.parent, st)); CUDA_CHECK(cudaFreeAsync(m->refit.boxes, st)); CUDA_CHECK(cudaFreeAsync(m->refit.counters, st));
        m->refit = RefitArrays{nullptr, nullptr, nullptr};
    }
    // The mesh keeps its triangle and node blocks when the triangle count is unchanged; everything transient lives in the device's
    // build arena.
    if (m->tris && n != m->n_tris) { CUDA_CHECK(cudaFreeAsync(m->tris, st)); m->tris = nullptr; }
    const uint32_t capacity = n ? wide_node_bound(n) : 0;
    if (m->nodes && m->node_capacity != capacity) { CUDA_CHECK(cudaFreeAsync(m->nodes, st)); m->nodes = nullptr; m->node_capacity = 0; }
    m->built = true; m->n_tris = n; m->generation++;
    m->n_nodes = m->n_packed = 0;
    memset(&m->stats, 0, sizeof(m->stats));
    m->stats.primitive_count = n;
    if (n == 0) return;
    std::lock_guard<std::mutex> arena_lock(d->build_mu);  // held while the build is enqueued, not while it runs
    const BuildScratch layout = build_scratch_layout(nullptr, n);
    if (!d->arena_event) CUDA_CHECK(cudaEventCreateWithFlags(&d->arena_event, cudaEventDisableTiming));
    else CUDA_CHECK(cudaStreamWaitEvent(st, d->arena_event, 0));  // the arena's previous user may be running on another stream
    if (layout.total_bytes > d->build_arena_cap) {
        CUDA_CHECK(cudaDeviceSynchronize());  // growing the arena: rare (sizes only grow), and nothing may still be using the old block
        if (d->build_arena) CUDA_CHECK(cudaFree(d->build_arena));
        d->build_arena_cap = layout.total_bytes + layout.total_bytes / 8;
        CUDA_CHECK(cudaMalloc((void **)&d->build_arena, d->build_arena_cap));
    }
    BuildScratch sc = build_scratch_layout(d->build_arena, n);
    // cubic curve BLASes keep the segments' power-basis coefficients behind the leaves, one 64-byte slot per segment (k_curve_coefs)
    const size_t leaf_slots = (size_t)n + (curve && curve->pieces > 1 ? n / curve->pieces : 0);
    if (!m->tris) CUDA_CHECK(cudaMallocAsync((void **)&m->tris, leaf_slots * sizeof(PackedTri), st));
    if (!m->nodes) { CUDA_CHECK(cudaMallocAsync((void **)&m->nodes, (size_t)capacity * sizeof(WideNode), st)); m->node_capacity = capacity; }
    // AccelUsageHint (api_types:204-212; the CPU backend maps it to Embree build quality, cpu/accel.rs:49-63): FastTrace (the default)
    // lets the builder choose between PLOC and the LBVH split rule per mesh, FastBuild always takes the LBVH.
    // LC_B200_BUILDER=lbvh|ploc|auto overrides.
    const int forced = g_builder_override.load();
    int builder = forced >= 0 ? forced : (m->option.hint == LCB_HINT_FAST_TRACE ? kBuilderAuto : kBuilderLbvh);
    // the per-mesh choice reads a statistic back from the device (one stream synchronisation, and PLOC reads its live cluster count
    // back between chunks of iterations); a rebuild of the same mesh with the same triangle count reuses the previous answer, and the
    // LBVH pipeline never touches the host
    const bool was_auto = builder == kBuilderAuto;
    if (was_auto && m->auto_builder >= 0 && m->auto_builder_n == n) builder = m->auto_builder;
    m->pend.begin(st);
    int built_with = kBuilderLbvh;
    if (curve) build_curves(st, n, *curve, sc, m->nodes, capacity, m->tris, d->lc);
    else if (aabbs) build_procedural(st, n, aabbs, sc, m->nodes, capacity, m->tris, d->lc);
    else built_with = build_blas(st, n, in, sc, m->nodes, capacity, m->tris, d->lc, builder);
    if (was_auto) { m->auto_builder = built_with; m->auto_builder_n = n; }
    m->pend_refit = false; m->pend_builder = built_with;
    m->pend.end(st, sc.header);
    CUDA_CHECK(cudaEventRecord(d->arena_event, st));
}

// ---- accel build (AccelImpl::update, cpu/accel.rs:324-447) ---------------------------------------
void invert_affine(const float m[12], float inv[12]) {
    // double adjugate inverse, one rounding to fp32 (restated independently in oracle.c)
    const double a = m[0], b = m[1], c = m[2], d = m[4], e = m[5], f = m[6], g = m[8], h = m[9], i = m[10];
    const double tx = m[3], ty = m[7], tz = m[11];
    const double c00 = e * i - f * h, c01 = c * h - b * i, c02 = b * f - c * e;
    const double c10 = f * g - d * i, c11 = a * i - c * g, c12 = c * d - a * f;
    const double c20 = d * h - e * g, c21 = b * g - a * h, c22 = a * e - b * d;
    const double det = a * c00 + b * c10 + c * c20;
    const double r = 1.0 / det;
    const double n00 = c00 * r, n01 = c01 * r, n02 = c02 * r, n10 = c10 * r, n11 = c11 * r, n12 = c12 * r, n20 = c20 * r, n21 = c21 * r, n22 = c22 * r;
    inv[0] = (float)n00; inv[1] = (float)n01; inv[2] = (float)n02; inv[3] = (float)(-(n00 * tx + n01 * ty + n02 * tz));
    inv[4] = (float)n10; inv[5] = (float)n11; inv[6] = (float)n12; inv[7] = (float)(-(n10 * tx + n11 * ty + n12 * tz));
    inv[8] = (float)n20; inv[9] = (float)n21; inv[10] = (float)n22; inv[11] = (float)(-(n20 * tx + n21 * ty + n22 * tz));
}

void accel_build(DeviceObj *d, StreamObj *s, const lcb_cmd_accel_build &c) {
    AccelObj *a = as<AccelObj>(c.accel.id);
    std::lock_guard<std::mutex> lk(a->mu);
    nvtx_range range("lc_b200 AccelBuild");
    cudaStream_t st = s->stream;
    accel_finalize(a);  // the previous build of this accel (its pinned staging block and record are reused)
    // Kernels may have edited the instance table (RayTracingSetInstance*): fold those edits into the host mirror first, so that this
    // build (and its TLAS) sees them and later modifications apply on top.  Only kernels that contain such a call raise maybe_dirty
    // (shader_dispatch), so ordinary frames — path tracers, ray queries — never wait for the stream here.
    if (a->maybe_dirty && a->dirty && a->table) {
        a->maybe_dirty = false;
        uint32_t flag = 0;
        CUDA_CHECK(cudaMemcpyAsync(&flag, a->dirty, 4, cudaMemcpyDeviceToHost, st));
        CUDA_CHECK(cudaStreamSynchronize(st));
        if (flag) {
            const size_t live = std::min<size_t>(a->instances.size(), a->table_capacity);
            std::vector<InstanceRec> recs(live);
            CUDA_CHECK(cudaMemcpyAsync(recs.data(), a->table, live * sizeof(InstanceRec), cudaMemcpyDeviceToHost, st));
            CUDA_CHECK(cudaMemsetAsync(a->dirty, 0, 4, st));
            CUDA_CHECK(cudaStreamSynchronize(st));
            for (size_t i = 0; i < live; i++) {
                InstanceHost &in = a->instances[i];
                if (!in.valid) continue;
                memcpy(in.affine, recs[i].affine, sizeof(in.affine));
                in.visible = recs[i].visibility; in.user_id = recs[i].user_id; in.opaque = (recs[i].flags & 2u) != 0;
            }
        }
    }
    a->pend.begin(st);
    const uint32_t n = c.instance_count;
    a->instances.resize(n);  // grow with default (invalid) slots / pop from the back (accel.rs:345-353)
    std::vector<uint8_t> touched(n, 0);
    for (size_t k = 0; k < c.modifications_count; k++) {
        const lcb_accel_modification &m = c.modifications[k];
        if (m.index >= n) fatal("AccelBuild: modification index %u out of range (%u instances)", m.index, n);
        InstanceHost &in = a->instances[m.index];
        touched[m.index] = 1;
        if (m.flags & LCB_MOD_PRIMITIVE) {  // accel.rs:355-377: resets mask/opaque, takes affine + user_id as given
            MeshObj *mesh = as<MeshObj>(m.mesh);
            if (!mesh->built) fatal("Mesh not built");
            memcpy(in.affine, m.affine, sizeof(in.affine));
            in.visible = 0xff; in.user_id = m.user_id; in.opaque = true; in.valid = true; in.mesh = mesh;
        }
        if (m.flags & LCB_MOD_OPAQUE_ON) in.opaque = true;
        else if (m.flags & LCB_MOD_OPAQUE_OFF) in.opaque = false;
        if (m.flags & LCB_MOD_TRANSFORM) { if (!in.valid) fatal("AccelBuild: TRANSFORM on an empty instance slot"); memcpy(in.affine, m.affine, sizeof(in.affine)); }
        if (m.flags & LCB_MOD_VISIBILITY) { if (!in.valid) fatal("AccelBuild: VISIBILITY on an empty instance slot"); in.visible = m.visibility; }
        if (m.flags & LCB_MOD_USER_ID) { if (!in.valid) fatal("AccelBuild: USER_ID on an empty instance slot"); in.user_id = m.user_id; }
    }
    // device table capacity
    if (n > a->table_capacity) {
        InstanceRec *nt = nullptr;
        uint32_t cap = n < 16 ? 16 : n + n / 2;
        CUDA_CHECK(cudaMallocAsync((void **)&nt, (size_t)cap * sizeof(InstanceRec), st));
        CUDA_CHECK(cudaMemsetAsync(nt, 0, (size_t)cap * sizeof(InstanceRec), st));
        if (a->table) {
            CUDA_CHECK(cudaMemcpyAsync(nt, a->table, (size_t)a->table_capacity * sizeof(InstanceRec), cudaMemcpyDeviceToDevice, st));
            CUDA_CHECK(cudaFreeAsync(a->table, st));
        }
        a->table = nt; a->table_capacity = cap;
    }
    // one record per touched slot, plus slots whose mesh was re-allocated since they were last resolved
    // (what update_accel_instance_handles does for OptiX, cuda_builtin_kernels.cu:42-50)
    std::vector<InstanceModRec> recs;
    std::vector<uint32_t> active;
    for (uint32_t i = 0; i < n; i++) {
        InstanceHost &in = a->instances[i];
        if (in.valid && in.mesh->generation != in.mesh_generation) touched[i] = 1;
        if (in.valid && in.mesh->n_tris > 0) active.push_back(i);
        if (!touched[i]) continue;
        InstanceModRec r{};
        r.index = i; r.visibility = in.visible; r.user_id = in.user_id;
        r.flags = (in.valid ? 1u : 0u) | (in.opaque ? 2u : 0u) | (in.valid && in.mesh->procedural ? 4u : 0u) | (in.valid && in.mesh->curve ? 8u : 0u);
        memcpy(r.affine, in.affine, sizeof(r.affine));
        invert_affine(in.affine, r.inv);
        bool identity = true;  // bit 4: the inverse is exactly the identity, zeros of either sign (trace_device.cuh enter_instance)
        for (int k = 0; k < 12; k++) identity = identity && r.inv[k] == ((k == 0 || k == 5 || k == 10) ? 1.0f : 0.0f);
        if (identity) r.flags |= 16u;
        if (in.valid) { r.nodes = in.mesh->nodes; r.tris = in.mesh->tris; in.mesh_generation = in.mesh->generation; }
        recs.push_back(r);
    }
    // one pinned staging block: [records | active ids]
    const size_t rec_bytes = recs.size() * sizeof(InstanceModRec), act_bytes = active.size() * 4;
    if (rec_bytes + act_bytes > a->h_stage_cap) {
        if (a->h_stage) cudaFreeHost(a->h_stage);
        a->h_stage_cap = (rec_bytes + act_bytes) * 2 + 4096;
        CUDA_CHECK(cudaMallocHost((void **)&a->h_stage, a->h_stage_cap));
    }
    if (!recs.empty()) {
        InstanceModRec *dv = nullptr;
        memcpy(a->h_stage, recs.data(), rec_bytes);
        CUDA_CHECK(cudaMallocAsync((void **)&dv, rec_bytes, st));
        CUDA_CHECK(cudaMemcpyAsync(dv, a->h_stage, rec_bytes, cudaMemcpyHostToDevice, st));
        apply_instance_mods(st, a->table, dv, (uint32_t)recs.size(), d->lc);
        CUDA_CHECK(cudaFreeAsync(dv, st));
    }
    a->stats.primitive_count = n;
    if (c.update_instance_buffer_only) {  // accel.rs:428-430
        a->pend_table_only = true;
        a->pend.end(st, nullptr);
        return;
    }
    // TLAS over the active instances (request is ignored, as on the CPU backend: stream.rs:418-428)
    const uint32_t na = (uint32_t)active.size();
    a->n_active = na;
    if (na > a->tlas_capacity) {
        if (a->tlas_nodes) { CUDA_CHECK(cudaFreeAsync(a->tlas_nodes, st)); CUDA_CHECK(cudaFreeAsync(a->tlas_prims, st)); CUDA_CHECK(cudaFreeAsync(a->active_ids, st)); }
        uint32_t cap = na < 16 ? 16 : na + na / 2;
        CUDA_CHECK(cudaMallocAsync((void **)&a->tlas_nodes, (size_t)cap * sizeof(WideNode), st));
        CUDA_CHECK(cudaMallocAsync((void **)&a->tlas_prims, (size_t)cap * 4, st));
        CUDA_CHECK(cudaMallocAsync((void **)&a->active_ids, (size_t)cap * 4, st));
        a->tlas_capacity = cap;
    }
    if (na) {
        memcpy(a->h_stage + rec_bytes, active.data(), act_bytes);
        CUDA_CHECK(cudaMemcpyAsync(a->active_ids, a->h_stage + rec_bytes, act_bytes, cudaMemcpyHostToDevice, st));
        BuildScratch layout = build_scratch_layout(nullptr, na);
        void *scratch = nullptr;
        CUDA_CHECK(cudaMallocAsync(&scratch, layout.total_bytes, st));
        BuildScratch sc = build_scratch_layout(scratch, na);
        build_tlas(st, na, a->active_ids, a->table, sc, a->tlas_nodes, a->tlas_prims, d->lc);
        a->pend_instances = n; a->pend_active = na; a->pend_table_only = false;
        a->pend.end(st, sc.header);
        CUDA_CHECK(cudaFreeAsync(scratch, st));
    } else {
        a->pend_instances = n; a->pend_active = 0; a->pend_table_only = false;
        a->pend.end(st, nullptr);
    }
}

// Fold a finished AccelBuild into the host-side record.  Caller holds a->mu.
void accel_finalize(AccelObj *a) {
    if (!a->pend.active) return;
    const float ms = a->pend.wait_ms();
    if (a->pend_table_only) return;
    const BuildHeader &hdr = *a->pend.h_hdr;
    if (a->pend_active) {
        if (hdr.error) fatal("TLAS build failed (code %u)", hdr.error);
        if (hdr.emitted != a->pend_active) fatal("TLAS build inconsistent: emitted %u of %u instances", hdr.emitted, a->pend_active);
        for (int k = 0; k < 3; k++) { a->world_lo[k] = hdr.root_lo[k]; a->world_hi[k] = hdr.root_hi[k]; }
        a->stats.wide_node_count = hdr.node_count; a->stats.packed_tri_count = hdr.prim_count; a->stats.max_depth = hdr.max_depth;
        a->stats.bvh_bytes = (uint64_t)hdr.node_count * sizeof(WideNode) + (uint64_t)a->pend_instances * sizeof(InstanceRec);
    } else {
        a->stats.wide_node_count = 0; a->stats.packed_tri_count = 0; a->stats.max_depth = 0; a->stats.bvh_bytes = 0;
    }
    a->stats.build_ms = ms; a->stats.was_refit = 0;
}

AccelView view_of(AccelObj *a) {
    if (a->pend.active && cudaEventQuery(a->pend.e1) == cudaSuccess) { std::lock_guard<std::mutex> lk(a->mu); accel_finalize(a); }  // world bounds, never blocks
    else (void)cudaGetLastError();
    AccelView v{};
    v.tlas_nodes = a->n_active ? a->tlas_nodes : nullptr;
    v.tlas_prims = a->tlas_prims; v.instances = a->table; v.instance_count = (uint32_t)a->instances.size();
    for (int k = 0; k < 3; k++) { v.world_lo[k] = a->world_lo[k]; v.world_hi[k] = a->world_hi[k]; }
    for (const InstanceHost &in : a->instances) if (in.valid && in.mesh->curve) { v.flags |= 1u; break; }
    return v;
}

size_t texture_region_bytes(const TextureObj *t, int32_t storage, uint32_t level, const uint32_t size[3], const char *what);
void shader_dispatch(DeviceObj *d, StreamObj *s, const lcb_cmd_shader_dispatch &c);
void bindless_update(StreamObj *s, const lcb_cmd_bindless_update &c);

// ---- dispatch (RustBackend::dispatch + StreamImpl::dispatch, cpu/mod.rs:168-180, stream.rs:213-446) ----
void dispatch(lcb_device dev, lcb_stream sh, lcb_command_list list, lcb_dispatch_callback cb, uint8_t *ctx) {
    DeviceObj *d = dev_of(dev); bind(d);
    StreamObj *s = as<StreamObj>(sh.id);
    cudaStream_t st = s->stream;
    std::vector<std::pair<void *, size_t>> staged;
    std::vector<cudaEvent_t> direct_copies;
    // Snapshot the source before returning (PinnedPool comment).  Two ways: (1) large pinned sources on an idle stream are copied by the
    // copy engine straight from the caller's memory and dispatch() waits for exactly that copy before it returns — the device buffer is
    // the snapshot, no second pass over host memory; (2) everything else goes through a pinned staging block.
    auto upload = [&](void *dst, const void *src, size_t n) {
        cudaPointerAttributes attr{};
        const bool pinned = n >= (size_t(1) << 20) && cudaPointerGetAttributes(&attr, src) == cudaSuccess && attr.type == cudaMemoryTypeHost;
        if (!pinned) (void)cudaGetLastError();
        // "idle" tolerates what drains in microseconds (the signal kernel of the previous chunk's timeline event): a short poll, not one query
        auto idle_soon = [&] {
            const auto t0 = std::chrono::steady_clock::now();
            for (;;) {
                if (cudaStreamQuery(st) == cudaSuccess) return true;
                if (std::chrono::steady_clock::now() - t0 > std::chrono::microseconds(100)) return false;
            }
        };
        if (pinned && (!direct_copies.empty() || idle_soon())) {
            CUDA_CHECK(cudaMemcpyAsync(dst, src, n, cudaMemcpyHostToDevice, st));
            cudaEvent_t ev; CUDA_CHECK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));  // spin-wait: a blocking-sync wake-up cost ~0.4 ms per 64 MB chunk (profiles/r02j_e2e_probe.txt)
            CUDA_CHECK(cudaEventRecord(ev, st));
            direct_copies.push_back(ev);
            return;
        }
        (void)cudaGetLastError();  // cudaStreamQuery: cudaErrorNotReady is not an error
        size_t cap = 0;
        void *stage = g_pinned.take(n, cap);
        parallel_copy(stage, src, n);
        staged.emplace_back(stage, cap);
        CUDA_CHECK(cudaMemcpyAsync(dst, stage, n, cudaMemcpyHostToDevice, st));
    };
    // Downloads never block the caller either: cudaMemcpyAsync into pageable memory would wait on the host until the copy has run — behind
    // everything queued on the stream, an event wait included (the reference only enqueues, cpu/mod.rs:168-180).  Pinned destinations are
    // written by the copy engine; pageable ones receive the bytes from a pinned staging block, copied by the stream's worker thread before
    // the completion callback fires (the frontend keeps download destinations alive until then, runtime.rs:1019-1027).
    std::vector<StreamObj::CopyOut> copy_out;
    auto download = [&](void *dst, const void *src, size_t n) {
        cudaPointerAttributes attr{};
        const bool pinned = cudaPointerGetAttributes(&attr, dst) == cudaSuccess && attr.type == cudaMemoryTypeHost;
        if (pinned) { CUDA_CHECK(cudaMemcpyAsync(dst, src, n, cudaMemcpyDeviceToHost, st)); return; }
        (void)cudaGetLastError();
        size_t cap = 0;
        void *stage = g_pinned.take(n, cap);
        CUDA_CHECK(cudaMemcpyAsync(stage, src, n, cudaMemcpyDeviceToHost, st));
        copy_out.push_back(StreamObj::CopyOut{dst, stage, n, cap});
    };
    for (size_t i = 0; i < list.commands_count; i++) {
        const lcb_command &c = list.commands[i];
        switch (c.tag) {
            case LCB_CMD_BUFFER_UPLOAD: {
                BufferObj *b = as<BufferObj>(c.u.buffer_upload.buffer.id);
                if (c.u.buffer_upload.offset + c.u.buffer_upload.size > b->size) fatal("BufferUpload out of range");
                if (c.u.buffer_upload.size) upload(b->ptr + c.u.buffer_upload.offset, c.u.buffer_upload.data, c.u.buffer_upload.size);
                break;
            }
            case LCB_CMD_BUFFER_DOWNLOAD: {
                BufferObj *b = as<BufferObj>(c.u.buffer_download.buffer.id);
                if (c.u.buffer_download.offset + c.u.buffer_download.size > b->size) fatal("BufferDownload out of range");
                if (c.u.buffer_download.size) download(c.u.buffer_download.data, b->ptr + c.u.buffer_download.offset, c.u.buffer_download.size);
                break;
            }
            case LCB_CMD_BUFFER_COPY: {
                BufferObj *src = as<BufferObj>(c.u.buffer_copy.src.id), *dst = as<BufferObj>(c.u.buffer_copy.dst.id);
                if (c.u.buffer_copy.src_offset + c.u.buffer_copy.size > src->size || c.u.buffer_copy.dst_offset + c.u.buffer_copy.size > dst->size) fatal("BufferCopy out of range");
                if (c.u.buffer_copy.size) CUDA_CHECK(cudaMemcpyAsync(dst->ptr + c.u.buffer_copy.dst_offset, src->ptr + c.u.buffer_copy.src_offset, c.u.buffer_copy.size, cudaMemcpyDeviceToDevice, st));
                break;
            }
            case LCB_CMD_TEXTURE_UPLOAD: {
                TextureObj *t = as<TextureObj>(c.u.texture_upload.texture.id);
                const size_t n = texture_region_bytes(t, c.u.texture_upload.storage, c.u.texture_upload.level, c.u.texture_upload.size, "TextureUpload");
                if (n) upload(t->ptr, c.u.texture_upload.data, n);
                break;
            }
            case LCB_CMD_TEXTURE_DOWNLOAD: {
                TextureObj *t = as<TextureObj>(c.u.texture_download.texture.id);
                const size_t n = texture_region_bytes(t, c.u.texture_download.storage, c.u.texture_download.level, c.u.texture_download.size, "TextureDownload");
                if (n) download(c.u.texture_download.data, t->ptr, n);
                break;
            }
            case LCB_CMD_TEXTURE_COPY: {
                TextureObj *src = as<TextureObj>(c.u.texture_copy.src.id), *dst = as<TextureObj>(c.u.texture_copy.dst.id);
                const size_t n = texture_region_bytes(src, c.u.texture_copy.storage, c.u.texture_copy.src_level, c.u.texture_copy.size, "TextureCopy(src)");
                if (texture_region_bytes(dst, c.u.texture_copy.storage, c.u.texture_copy.dst_level, c.u.texture_copy.size, "TextureCopy(dst)") != n) fatal("TextureCopy: size mismatch");
                if (n) CUDA_CHECK(cudaMemcpyAsync(dst->ptr, src->ptr, n, cudaMemcpyDeviceToDevice, st));
                break;
            }
            case LCB_CMD_BUFFER_TO_TEXTURE: case LCB_CMD_TEXTURE_TO_BUFFER: {
                const lcb_cmd_buffer_texture &bt = c.tag == LCB_CMD_BUFFER_TO_TEXTURE ? c.u.buffer_to_texture : c.u.texture_to_buffer;
                BufferObj *b = as<BufferObj>(bt.buffer.id); TextureObj *t = as<TextureObj>(bt.texture.id);
                const size_t n = texture_region_bytes(t, bt.storage, bt.level, bt.size, "Buffer<->Texture copy");
                if (bt.buffer_offset + n > b->size) fatal("Buffer<->Texture copy: buffer range out of bounds");
                if (n) {
                    if (c.tag == LCB_CMD_BUFFER_TO_TEXTURE) CUDA_CHECK(cudaMemcpyAsync(t->ptr, b->ptr + bt.buffer_offset, n, cudaMemcpyDeviceToDevice, st));
                    else CUDA_CHECK(cudaMemcpyAsync(b->ptr + bt.buffer_offset, t->ptr, n, cudaMemcpyDeviceToDevice, st));
                }
                break;
            }
            case LCB_CMD_SHADER_DISPATCH: shader_dispatch(d, s, c.u.shader_dispatch); break;
            case LCB_CMD_BINDLESS_UPDATE: bindless_update(s, c.u.bindless_update); break;
            case LCB_CMD_MESH_BUILD: mesh_build(d, s, c.u.mesh_build); break;
            case LCB_CMD_PROCEDURAL_BUILD: procedural_build(d, s, c.u.procedural_build); break;
            case LCB_CMD_CURVE_BUILD: curve_build(d, s, c.u.curve_build); break;
            case LCB_CMD_ACCEL_BUILD: accel_build(d, s, c.u.accel_build); break;
            default:
                fatal("command tag %d is outside the B200 ray-tracing device's scope (SURVEY.md §8)", c.tag);
        }
    }
    flush_launches(d);
    StreamObj::Pending p{};
    CUDA_CHECK(cudaEventCreateWithFlags(&p.ev, cudaEventDisableTiming));
    CUDA_CHECK(cudaEventRecord(p.ev, st));
    p.cb = cb; p.ctx = ctx; p.staged = std::move(staged); p.copy_out = std::move(copy_out);
    s->push(std::move(p));
    for (cudaEvent_t ev : direct_copies) { CUDA_CHECK(cudaEventSynchronize(ev)); cudaEventDestroy(ev); }  // the borrow of those sources ends here
}

// ---- events (timeline semantics, cpu/resource.rs:10-44, cpu/mod.rs:367-402) -----------------------
__global__ void k_event_signal(unsigned long long *counter, unsigned long long value) { atomicMax(counter, value); }

typedef int (*StreamWaitValue64Fn)(cudaStream_t, unsigned long long /* CUdeviceptr */, unsigned long long, unsigned int);
StreamWaitValue64Fn stream_wait_value64() {
    static StreamWaitValue64Fn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult status;
        if (cudaGetDriverEntryPoint("cuStreamWaitValue64", &p, cudaEnableDefault, &status) != cudaSuccess || status != cudaDriverEntryPointSuccess || !p)
            fatal("the driver does not export cuStreamWaitValue64: timeline events need stream memory operations");
        return (StreamWaitValue64Fn)p;
    }();
    return fn;
}

lcb_created create_event(lcb_device dev) {
    bind(dev_of(dev));
    auto *e = new EventObj;
    CUDA_CHECK(cudaMalloc((void **)&e->counter, 256));
    zero_fill(e->counter, 256);
    return lcb_created{(uint64_t)e, e};
}
void destroy_event(lcb_device dev, lcb_event h) {
    bind(dev_of(dev));
    EventObj *e = as<EventObj>(h.id);
    cudaFree(e->counter);  // synchronises with every stream that still signals or waits on it
    delete e;
}
void signal_event(lcb_device dev, lcb_event h, lcb_stream sh, uint64_t value) {
    DeviceObj *d = dev_of(dev); bind(d);
    EventObj *e = as<EventObj>(h.id); StreamObj *s = as<StreamObj>(sh.id);
    k_event_signal<<<1, 1, 0, s->stream>>>(e->counter, (unsigned long long)value);
    CUDA_CHECK(cudaGetLastError());
    g_launches++;
}
uint64_t event_value(DeviceObj *d, EventObj *e) {  // the counter as the device holds it now
    std::lock_guard<std::mutex> lk(d->poll_mu);
    if (!d->poll_stream) {
        CUDA_CHECK(cudaStreamCreateWithFlags(&d->poll_stream, cudaStreamNonBlocking));
        CUDA_CHECK(cudaHostAlloc((void **)&d->poll_word, 64, cudaHostAllocDefault));
    }
    CUDA_CHECK(cudaMemcpyAsync(d->poll_word, e->counter, sizeof(unsigned long long), cudaMemcpyDeviceToHost, d->poll_stream));
    CUDA_CHECK(cudaStreamSynchronize(d->poll_stream));
    const uint64_t v = *d->poll_word;
    uint64_t prev = e->seen.load();
    while (prev < v && !e->seen.compare_exchange_weak(prev, v)) {}
    return v;
}
bool is_event_completed(lcb_device dev, lcb_event h, uint64_t value) {
    DeviceObj *d = dev_of(dev); bind(d);
    EventObj *e = as<EventObj>(h.id);
    return e->seen.load() >= value || event_value(d, e) >= value;
}
void synchronize_event(lcb_device dev, lcb_event h, uint64_t value) {
    DeviceObj *d = dev_of(dev); bind(d);
    EventObj *e = as<EventObj>(h.id);
    if (e->seen.load() >= value) return;
    for (unsigned spins = 0; event_value(d, e) < value; spins++) {
        if (spins > 64) std::this_thread::sleep_for(std::chrono::microseconds(spins > 1024 ? 200 : 20));
    }
}
void wait_event(lcb_device dev, lcb_event h, lcb_stream sh, uint64_t value) {
    bind(dev_of(dev));
    EventObj *e = as<EventObj>(h.id);
    if (e->seen.load() >= value) return;  // already completed: nothing to order against
    const int rc = stream_wait_value64()(as<StreamObj>(sh.id)->stream, (unsigned long long)e->counter, (unsigned long long)value, 0x0 /* CU_STREAM_WAIT_VALUE_GEQ */);
    if (rc != 0) fatal("cuStreamWaitValue64 failed (%d)", rc);
}

// ---- mesh / accel objects ------------------------------------------------------------------------
lcb_created create_mesh(lcb_device dev, const lcb_accel_option *opt) { bind(dev_of(dev)); auto *m = new MeshObj; if (opt) m->option = *opt; return lcb_created{(uint64_t)m, m}; }
void destroy_mesh(lcb_device dev, lcb_mesh h) {
    bind(dev_of(dev));
    MeshObj *m = as<MeshObj>(h.id);
    if (m->pend.active) m->pend.wait_ms();
    m->pend.destroy();
    if (m->nodes) cudaFree(m->nodes);
    if (m->tris) cudaFree(m->tris);
    if (m->refit.parent) { cudaFree(m->refit.parent); cudaFree(m->refit.boxes); cudaFree(m->refit.counters); }
    delete m;
}
lcb_created create_accel(lcb_device dev, const lcb_accel_option *opt) { bind(dev_of(dev)); auto *a = new AccelObj; if (opt) a->option = *opt; return lcb_created{(uint64_t)a, a}; }
void destroy_accel(lcb_device dev, lcb_accel h) {
    bind(dev_of(dev));
    AccelObj *a = as<AccelObj>(h.id);
    if (a->pend.active) a->pend.wait_ms();
    a->pend.destroy();
    if (a->table) cudaFree(a->table);
    if (a->tlas_nodes) { cudaFree(a->tlas_nodes); cudaFree(a->tlas_prims); cudaFree(a->active_ids); }
    if (a->h_stage) cudaFreeHost(a->h_stage);
    if (a->dirty) cudaFree(a->dirty);
    delete a;
}

// ---- textures (cpu/texture.rs, cpu/mod.rs:85-120) ---------------------------------------------------------------------
int32_t format_to_storage(int32_t format) {  // PixelFormat::storage, api_types:345-385
    static const int8_t map[30] = {0, 0, 0, 1, 1, 1, 2, 2, 2, 3, 3, 3, 4, 4, 4, 5, 5, 5, 6, 6, 7, 7, 8, 8, 9, 10, 11, 12, 13, 14};
    if (format < 0 || format >= 30) fatal("pixel format %d (packed / block-compressed) is not supported by the B200 device", format);
    return map[format];
}
size_t storage_pixel_bytes(int32_t storage) {  // PixelStorage::size, api_types:285-306
    static const size_t sz[15] = {1, 2, 4, 2, 4, 8, 4, 8, 16, 2, 4, 8, 4, 8, 16};
    if (storage < 0 || storage >= 15) fatal("pixel storage %d is not supported by the B200 device", storage);
    return sz[storage];
}
lcb_created create_texture(lcb_device dev, int32_t format, uint32_t dim, uint32_t w, uint32_t h, uint32_t d, uint32_t mips, bool, bool) {
    DeviceObj *dv = dev_of(dev); bind(dv);
    if (dim != 2 && dim != 3) fatal("create_texture: dimension must be 2 or 3 (got %u)", dim);
    if (mips > 1) fatal("create_texture: mipmapped textures are outside the ray-tracing device's scope (got %u levels)", mips);
    auto *t = new TextureObj;
    t->dim = dim; t->width = w; t->height = h; t->depth = dim == 2 ? 1 : d;
    t->storage = format_to_storage(format); t->pixel_bytes = storage_pixel_bytes(t->storage);
    t->bytes = (size_t)t->width * t->height * t->depth * t->pixel_bytes;
    CUDA_CHECK(cudaMalloc(&t->ptr, t->bytes ? t->bytes : 16));
    zero_fill(t->ptr, t->bytes ? t->bytes : 16);
    return lcb_created{(uint64_t)t, t->ptr};
}
void destroy_texture(lcb_device dev, lcb_texture h) { bind(dev_of(dev)); TextureObj *t = as<TextureObj>(h.id); cudaFree(t->ptr); delete t; }
size_t texture_region_bytes(const TextureObj *t, int32_t storage, uint32_t level, const uint32_t size[3], const char *what) {
    if (level != 0) fatal("%s: mip level %u of a single-level texture", what, level);
    if (storage != t->storage) fatal("%s: storage %d does not match the texture's storage %d", what, storage, t->storage);
    if (size[0] != t->width || size[1] != t->height || (t->dim == 3 && size[2] != t->depth)) fatal("%s: region %ux%ux%u is not the whole level (%ux%ux%u)", what, size[0], size[1], size[2], t->width, t->height, t->depth);
    return t->bytes;
}
HostTextureArg texture_arg(const TextureObj *t, uint32_t sampler = 0) { return HostTextureArg{t->ptr, t->width, t->height, t->depth, (uint32_t)t->storage, sampler, 0u}; }

// ---- bindless arrays -------------------------------------------------------------------------------------------------------
lcb_created create_bindless_array(lcb_device dev, size_t size) {
    bind(dev_of(dev));
    auto *b = new BindlessObj;
    b->host.assign(size, HostBindlessSlot{});
    CUDA_CHECK(cudaMalloc((void **)&b->device, (size ? size : 1) * sizeof(HostBindlessSlot)));
    zero_fill(b->device, (size ? size : 1) * sizeof(HostBindlessSlot));
    return lcb_created{(uint64_t)b, b->device};
}
void destroy_bindless_array(lcb_device dev, lcb_bindless h) { bind(dev_of(dev)); BindlessObj *b = as<BindlessObj>(h.id); cudaFree(b->device); delete b; }
void bindless_update(StreamObj *s, const lcb_cmd_bindless_update &c) {  // BindlessArrayImpl::update, cpu/resource.rs:76-124
    BindlessObj *b = as<BindlessObj>(c.handle.id);
    std::lock_guard<std::mutex> lk(b->mu);
    size_t lo = SIZE_MAX, hi = 0;
    for (size_t i = 0; i < c.modifications_count; i++) {
        const lcb_bindless_modification &m = c.modifications[i];
        if (m.slot >= b->host.size()) fatal("BindlessArrayUpdate: slot %zu out of range (%zu slots)", m.slot, b->host.size());
        HostBindlessSlot &slot = b->host[m.slot];
        if (m.buffer.op == 1) {
            BufferObj *buf = as<BufferObj>(m.buffer.handle.id);
            if (m.buffer.offset > buf->size) fatal("BindlessArrayUpdate: buffer offset beyond the buffer");
            slot.buffer = buf->ptr + m.buffer.offset; slot.buffer_size = buf->size - m.buffer.offset;
        } else if (m.buffer.op == 2) { slot.buffer = nullptr; slot.buffer_size = 0; }
        if (m.tex2d.op == 1) slot.tex2d = texture_arg(as<TextureObj>(m.tex2d.handle.id), (uint32_t)(m.tex2d.sampler.filter & 3) | ((uint32_t)(m.tex2d.sampler.address & 3) << 2)); else if (m.tex2d.op == 2) slot.tex2d = HostTextureArg{};
        if (m.tex3d.op == 1) slot.tex3d = texture_arg(as<TextureObj>(m.tex3d.handle.id), (uint32_t)(m.tex3d.sampler.filter & 3) | ((uint32_t)(m.tex3d.sampler.address & 3) << 2)); else if (m.tex3d.op == 2) slot.tex3d = HostTextureArg{};
        lo = std::min(lo, m.slot); hi = std::max(hi, m.slot + 1);
    }
    if (lo < hi)  // pageable source: staged before the call returns, the host table may change right after
        CUDA_CHECK(cudaMemcpyAsync(b->device + lo, b->host.data() + lo, (hi - lo) * sizeof(HostBindlessSlot), cudaMemcpyHostToDevice, s->stream));
}

// ---- shaders (ShaderImpl, cpu/shader.rs; dispatch cpu/stream.rs:330-440) ------------------------------------------------------
lcb_created_shader create_shader(lcb_device dev, lcb_kernel_module km, const lcb_shader_option *opt) {
    DeviceObj *d = dev_of(dev); bind(d);
    std::string log;
    ShaderObj *s = nullptr;
    try {
        s = shader_create(reinterpret_cast<const ir::KernelModule *>(km.ptr), opt && opt->enable_fast_math, opt && opt->compile_only, opt ? opt->name : nullptr, log);
    } catch (const std::exception &e) { fatal("create_shader: %s", e.what()); }
    if (!log.empty()) log_msg("W", "create_shader: NVRTC log:\n%s", log.c_str());
    lcb_created_shader out{};
    out.resource.handle = (uint64_t)s; out.resource.native_handle = s;
    for (int k = 0; k < 3; k++) out.block_size[k] = shader_lowered(s).block_size[k];
    return out;
}
void destroy_shader(lcb_device dev, lcb_shader h) { bind(dev_of(dev)); shader_destroy(as<ShaderObj>(h.id)); }

AccelView view_of(AccelObj *a);
void shader_dispatch(DeviceObj *d, StreamObj *s, const lcb_cmd_shader_dispatch &c) {
    ShaderObj *sh = as<ShaderObj>(c.shader.id);
    const LoweredKernel &k = shader_lowered(sh);
    if (c.args_count != k.args.size()) fatal("ShaderDispatch: %zu arguments given, the kernel takes %zu (runtime.rs:1517)", c.args_count, k.args.size());
    std::vector<uint8_t> block(k.param_bytes, 0);
    HostLaunch launch{{c.dispatch_size[0], c.dispatch_size[1], c.dispatch_size[2]}, 0, nullptr, 0};
    memcpy(block.data(), &launch, sizeof(launch));
    auto put_buffer = [&](const ParamSlot &p, uint64_t handle, size_t offset, size_t size) {
        BufferObj *b = as<BufferObj>(handle);
        if (offset + size > b->size) fatal("ShaderDispatch: buffer view [%zu, +%zu) exceeds the buffer (%zu bytes)", offset, size, b->size);
        HostBufferArg a{b->ptr + offset, size}; memcpy(block.data() + p.offset, &a, sizeof(a));
    };
    auto put_texture = [&](const ParamSlot &p, uint64_t handle, uint32_t level) {
        if (level != 0) fatal("ShaderDispatch: texture level %u of a single-level texture", level);
        HostTextureArg a = texture_arg(as<TextureObj>(handle)); memcpy(block.data() + p.offset, &a, sizeof(a));
    };
    auto put_bindless = [&](const ParamSlot &p, uint64_t handle) {
        BindlessObj *b = as<BindlessObj>(handle);
        HostBindlessArg a{b->device, b->host.size()}; memcpy(block.data() + p.offset, &a, sizeof(a));
    };
    auto put_accel = [&](const ParamSlot &p, uint64_t handle) {
        AccelObj *ao = as<AccelObj>(handle);
        if (!ao->dirty) { CUDA_CHECK(cudaMalloc((void **)&ao->dirty, 256)); zero_fill(ao->dirty, 256); }
        if (k.writes_accel) ao->maybe_dirty = true;
        HostAccelArg a{view_of(ao), ao->table, ao->dirty}; memcpy(block.data() + p.offset, &a, sizeof(a));
    };
    for (const ParamSlot &p : k.captures) {  // bound at create_shader time (KernelModule.captures, cpu/mod.rs:296-301)
        switch (p.kind) {
            case ParamSlot::Buffer: put_buffer(p, p.binding.buffer.handle, p.binding.buffer.offset, p.binding.buffer.size); break;
            case ParamSlot::Texture: put_texture(p, p.binding.texture.handle, p.binding.texture.level); break;
            case ParamSlot::Bindless: put_bindless(p, p.binding.bindless_array); break;
            case ParamSlot::Accel: put_accel(p, p.binding.accel); break;
            default: fatal("ShaderDispatch: bad capture kind");
        }
    }
    for (size_t i = 0; i < k.args.size(); i++) {
        const ParamSlot &p = k.args[i];
        const lcb_argument &a = c.args[i];
        static const int want[5] = {LCB_ARG_BUFFER, LCB_ARG_TEXTURE, LCB_ARG_BINDLESS, LCB_ARG_ACCEL, LCB_ARG_UNIFORM};
        if (a.tag != want[p.kind]) fatal("ShaderDispatch: argument %zu has tag %d, the kernel expects %d", i, a.tag, want[p.kind]);
        switch (p.kind) {
            case ParamSlot::Buffer: put_buffer(p, a.u.buffer.buffer.id, a.u.buffer.offset, a.u.buffer.size); break;
            case ParamSlot::Texture: put_texture(p, a.u.texture.texture.id, a.u.texture.level); break;
            case ParamSlot::Bindless: put_bindless(p, a.u.bindless.id); break;
            case ParamSlot::Accel: put_accel(p, a.u.accel.id); break;
            case ParamSlot::Uniform:
                if (a.u.uniform.size != p.size) fatal("ShaderDispatch: uniform %zu has %zu bytes, the kernel expects %zu", i, a.u.uniform.size, p.size);
                memcpy(block.data() + p.offset, a.u.uniform.data, p.size);
                break;
        }
    }
    try { d->lc.count += shader_launch(sh, s->stream, block.data(), c.dispatch_size, s->work_counter); } catch (const std::exception &e) { fatal("ShaderDispatch: %s", e.what()); }
}

// ---- out-of-scope slots: loud failure -------------------------------------------------------------
#define UNSUPPORTED(what) fatal("%s is outside the B200 ray-tracing device's scope (SURVEY.md §8: hot path only; no fallback)", what)
lcb_created_swapchain create_swapchain(lcb_device, const lcb_swapchain_option *, lcb_stream) { UNSUPPORTED("create_swapchain"); }
void present_display_in_stream(lcb_device, lcb_stream, lcb_swapchain, lcb_texture) { UNSUPPORTED("present_display_in_stream"); }
void destroy_swapchain(lcb_device, lcb_swapchain) { UNSUPPORTED("destroy_swapchain"); }
lcb_created create_curve(lcb_device dev, const lcb_accel_option *opt) {
    lcb_created c = create_mesh(dev, opt);
    as<MeshObj>(c.handle)->curve = true;
    return c;
}
void destroy_curve(lcb_device dev, lcb_curve h) { destroy_mesh(dev, lcb_mesh{h.id}); }
lcb_created create_procedural_primitive(lcb_device dev, const lcb_accel_option *opt) {
    bind(dev_of(dev));
    auto *m = new MeshObj; if (opt) m->option = *opt;
    m->procedural = true;
    return lcb_created{(uint64_t)m, m};
}
void destroy_procedural_primitive(lcb_device dev, lcb_procedural h) { destroy_mesh(dev, lcb_mesh{h.id}); }

void *native_handle(lcb_device dev) { return dev_of(dev); }
uint32_t compute_warp_size(lcb_device) { return 32; }
char *query(lcb_device, const char *name) {
    // backend/lib.rs:447-457: "" reads as None on the Rust side
    const char *v = "";
    if (name && strcmp(name, "device_name") == 0) v = "b200";
    char *out = (char *)malloc(strlen(v) + 1); strcpy(out, v); return out;
}
lcb_pinned_memory_ext pinned_memory_ext(lcb_device) { return lcb_pinned_memory_ext{nullptr, nullptr, nullptr}; }
lcb_denoiser_ext denoiser_ext(lcb_device) { return lcb_denoiser_ext{nullptr, nullptr, nullptr, nullptr, nullptr}; }  // data == nullptr => invalid (api_types:1020-1024)

void destroy_device(lcb_device_interface iface) {
    DeviceObj *d = dev_of(iface.device); bind(d);
    if (d->internal) free_stream(d->internal);
    if (d->internal2) free_stream(d->internal2);
    if (d->copy_in) cudaStreamDestroy(d->copy_in);
    if (d->copy_out) cudaStreamDestroy(d->copy_out);
    if (d->stage_rays) cudaFree(d->stage_rays);
    if (d->stage_out) cudaFree(d->stage_out);
    if (d->build_arena) cudaFree(d->build_arena);
    if (d->poll_stream) cudaStreamDestroy(d->poll_stream);
    if (d->poll_word) cudaFreeHost(d->poll_word);
    g_pinned.clear();
    flush_launches(d);
    delete d;
}

// Device names this library does not own are forwarded to the stock backend library when one is installed next to it: the frontend
// dlopens a fixed file name (liblc-api.so, luisa_compute_backend/src/lib.rs:102-117), so a drop-in that keeps "cpu" / "cuda" working is
// this library under that name plus the original renamed to liblc-api-orig.so (or named by LC_B200_FORWARD_LIB).  SURVEY.md §8b.
struct Forward { void *handle = nullptr; lcb_lib_interface iface{}; lcb_context ctx{0}; bool tried = false; std::mutex mu; };
Forward g_forward;

bool forward_ready(const char *runtime_dir) {
    std::lock_guard<std::mutex> lk(g_forward.mu);
    if (g_forward.tried) return g_forward.handle != nullptr;
    g_forward.tried = true;
    std::vector<std::string> candidates;
    if (const char *env = getenv("LC_B200_FORWARD_LIB")) candidates.push_back(env);
    Dl_info info;
    if (dladdr((const void *)&forward_ready, &info) && info.dli_fname) {
        std::string dir = info.dli_fname;
        const size_t slash = dir.rfind('/');
        dir = slash == std::string::npos ? "." : dir.substr(0, slash);
        candidates.push_back(dir + "/liblc-api-orig.so");
    }
    for (const std::string &c : candidates) {
        void *h = dlopen(c.c_str(), RTLD_NOW | RTLD_LOCAL);
        if (!h) continue;
        auto entry = (lcb_lib_interface(*)(void))dlsym(h, "luisa_compute_lib_interface");
        if (!entry) { dlclose(h); continue; }
        g_forward.handle = h;
        g_forward.iface = entry();
        if (g_forward.iface.set_logger_callback && g_logger.load()) g_forward.iface.set_logger_callback(g_logger.load());
        g_forward.ctx = g_forward.iface.create_context(runtime_dir ? runtime_dir : ".");
        log_msg("I", "device names other than \"b200\" are forwarded to %s", c.c_str());
        return true;
    }
    return false;
}

// ---- library interface -----------------------------------------------------------------------------
void set_logger_callback(void (*cb)(lcb_logger_message)) {
    g_logger.store(cb);
    std::lock_guard<std::mutex> lk(g_forward.mu);
    if (g_forward.handle && g_forward.iface.set_logger_callback) g_forward.iface.set_logger_callback(cb);
}
lcb_context create_context(const char *) { return lcb_context{1}; }
void destroy_context(lcb_context) {}
void free_string(char *s) { free(s); }

lcb_device_interface create_device(lcb_context, const char *name, const char *json) {
    if (!name || (strcmp(name, "b200") != 0 && strcmp(name, "cuda-b200") != 0)) {
        if (name && forward_ready(nullptr)) return g_forward.iface.create_device(g_forward.ctx, name, json);
        fatal("device \"%s\" is not served by this library (only \"b200\") and no stock backend library (liblc-api-orig.so / LC_B200_FORWARD_LIB) "
              "is installed to forward to; there is no CPU fallback", name ? name : "(null)");
    }
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) fatal("no CUDA device available (%s); the b200 device has no CPU fallback", cudaGetErrorString(e));
    int ordinal = 0;
    if (const char *env = getenv("LC_B200_DEVICE")) ordinal = atoi(env);
    else if (const char *lr = getenv("LOCAL_RANK")) ordinal = atoi(lr) % count;
    if (json) { const char *p = strstr(json, "\"device_index\""); if (p && (p = strchr(p, ':'))) ordinal = atoi(p + 1); }
    if (ordinal < 0 || ordinal >= count) fatal("device index %d out of range (%d devices)", ordinal, count);
    CUDA_CHECK(cudaSetDevice(ordinal));
    cudaDeviceProp prop; CUDA_CHECK(cudaGetDeviceProperties(&prop, ordinal));
    if (prop.major != 10) fatal("device %d is sm_%d%d; this library contains sm_100a code only", ordinal, prop.major, prop.minor);
    {   // builds allocate scratch + result arrays stream-ordered; keep freed blocks in the pool instead of returning
        // them to the driver at every synchronisation (the default threshold of 0 makes each rebuild re-map memory)
        cudaMemPool_t pool; CUDA_CHECK(cudaDeviceGetDefaultMemPool(&pool, ordinal));
        unsigned long long keep = ~0ull; CUDA_CHECK(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
    }
    auto *d = new DeviceObj; d->ordinal = ordinal;
    d->internal = make_stream(d); d->internal2 = make_stream(d);
    CUDA_CHECK(cudaStreamCreateWithFlags(&d->copy_in, cudaStreamNonBlocking));
    CUDA_CHECK(cudaStreamCreateWithFlags(&d->copy_out, cudaStreamNonBlocking));
    log_msg("I", "b200 device %d: %s, %d SMs, %.1f GB", ordinal, prop.name, prop.multiProcessorCount, prop.totalGlobalMem / 1e9);
    lcb_device_interface t{};
    t.device = lcb_device{(uint64_t)d};
    t.destroy_device = destroy_device; t.create_buffer = create_buffer; t.destroy_buffer = destroy_buffer;
    t.create_texture = create_texture; t.native_handle = native_handle; t.compute_warp_size = compute_warp_size;
    t.destroy_texture = destroy_texture; t.create_bindless_array = create_bindless_array; t.destroy_bindless_array = destroy_bindless_array;
    t.create_stream = create_stream; t.destroy_stream = destroy_stream; t.synchronize_stream = synchronize_stream; t.dispatch = dispatch;
    t.create_swapchain = create_swapchain; t.present_display_in_stream = present_display_in_stream; t.destroy_swapchain = destroy_swapchain;
    t.create_shader = create_shader; t.destroy_shader = destroy_shader;
    t.create_event = create_event; t.destroy_event = destroy_event; t.signal_event = signal_event; t.synchronize_event = synchronize_event;
    t.wait_event = wait_event; t.is_event_completed = is_event_completed;
    t.create_mesh = create_mesh; t.destroy_mesh = destroy_mesh; t.create_curve = create_curve; t.destroy_curve = destroy_curve;
    t.create_procedural_primitive = create_procedural_primitive; t.destroy_procedural_primitive = destroy_procedural_primitive;
    t.create_accel = create_accel; t.destroy_accel = destroy_accel; t.query = query;
    t.pinned_memory_ext = pinned_memory_ext; t.denoiser_ext = denoiser_ext;
    return t;
}

void ensure_stage(uint8_t *&p, size_t &cap, size_t need) {
    if (need <= cap) return;
    if (p) CUDA_CHECK(cudaFree(p));
    cap = need + need / 4;
    CUDA_CHECK(cudaMalloc(&p, cap));
}

}  // namespace

// ================================= exported symbols ===================================================

extern "C" {

lcb_lib_interface luisa_compute_lib_interface(void) {
    lcb_lib_interface l{};
    l.inner = nullptr; l.set_logger_callback = set_logger_callback; l.create_context = create_context;
    l.destroy_context = destroy_context; l.create_device = create_device; l.free_string = free_string;
    return l;
}

void lc_b200_trace_closest(lcb_device dev, lcb_stream sh, lcb_accel ah, lcb_buffer rays, size_t rays_offset, lcb_buffer hits, size_t hits_offset,
                           uint64_t count, uint32_t mask) {
    DeviceObj *d = dev_of(dev); bind(d);
    StreamObj *s = as<StreamObj>(sh.id); AccelObj *a = as<AccelObj>(ah.id);
    BufferObj *rb = as<BufferObj>(rays.id), *hb = as<BufferObj>(hits.id);
    if (rays_offset % 16 || hits_offset % 8) fatal("trace_closest: misaligned offsets");
    if (rays_offset + count * 32 > rb->size || hits_offset + count * 24 > hb->size) fatal("trace_closest: %llu rays exceed the buffers", (unsigned long long)count);
    if (count) trace_closest(s->stream, view_of(a), rb->ptr + rays_offset, hb->ptr + hits_offset, count, mask, s->work_counter, nullptr, d->lc);
    flush_launches(d);
}

void lc_b200_trace_any(lcb_device dev, lcb_stream sh, lcb_accel ah, lcb_buffer rays, size_t rays_offset, lcb_buffer occ, size_t occ_offset, uint64_t count,
                       uint32_t mask) {
    DeviceObj *d = dev_of(dev); bind(d);
    StreamObj *s = as<StreamObj>(sh.id); AccelObj *a = as<AccelObj>(ah.id);
    BufferObj *rb = as<BufferObj>(rays.id), *ob = as<BufferObj>(occ.id);
    if (rays_offset % 16 || occ_offset % 4) fatal("trace_any: misaligned offsets");
    if (rays_offset + count * 32 > rb->size || occ_offset + count * 4 > ob->size) fatal("trace_any: %llu rays exceed the buffers", (unsigned long long)count);
    if (count) trace_any(s->stream, view_of(a), rb->ptr + rays_offset, (uint32_t *)(ob->ptr + occ_offset), count, mask, s->work_counter, d->lc);
    flush_launches(d);
}

void lc_b200_ray_query(lcb_device dev, lcb_stream sh, lcb_accel ah, lcb_buffer rays, size_t rays_offset, lcb_buffer committed, size_t committed_offset,
                       uint64_t count, uint32_t mask, bool terminate_on_first, const lcb_candidate_filter *filter) {
    DeviceObj *d = dev_of(dev); bind(d);
    StreamObj *s = as<StreamObj>(sh.id); AccelObj *a = as<AccelObj>(ah.id);
    BufferObj *rb = as<BufferObj>(rays.id), *hb = as<BufferObj>(committed.id);
    if (rays_offset % 16 || committed_offset % 8) fatal("ray_query: misaligned offsets");
    if (rays_offset + count * 32 > rb->size || committed_offset + count * 24 > hb->size) fatal("ray_query: %llu rays exceed the buffers", (unsigned long long)count);
    CandidateFilter f{LCB_FILTER_COMMIT_ALL, 0.f, nullptr, nullptr};
    if (filter) {
        if (filter->kind < 0 || filter->kind > LCB_FILTER_REJECT_ALL) fatal("ray_query: unknown candidate filter %d", filter->kind);
        f.kind = filter->kind; f.radius = filter->radius;
        if (filter->kind == LCB_FILTER_PRIM_BITS) {
            BufferObj *bits = as<BufferObj>(filter->bits.id), *first = as<BufferObj>(filter->first_bit.id);
            if (first->size < a->instances.size() * 4) fatal("ray_query: first_bit needs one uint32 per instance slot");
            f.bits = (const uint32_t *)bits->ptr; f.first_bit = (const uint32_t *)first->ptr;
        }
    }
    if (count) ray_query(s->stream, view_of(a), rb->ptr + rays_offset, hb->ptr + committed_offset, count, mask, terminate_on_first, f, s->work_counter, d->lc);
    flush_launches(d);
}

void lc_b200_example_path_tracer(lcb_device dev, lcb_stream sh, const lcb_path_tracer_args *args, uint64_t ray_counts_out[2]) {
    DeviceObj *d = dev_of(dev); bind(d);
    StreamObj *s = as<StreamObj>(sh.id); AccelObj *a = as<AccelObj>(args->accel.id);
    BufferObj *img = as<BufferObj>(args->image.id), *seed = as<BufferObj>(args->seed_image.id);
    const size_t px = (size_t)args->width * args->height;
    if (img->size < px * 16 || seed->size < px * 4) fatal("path_tracer: image / seed buffers are smaller than %u x %u", args->width, args->height);
    if (args->heap_size < a->instances.size()) fatal("path_tracer: heaps must cover every instance slot");
    // device-side pointer tables of the two bindless heaps + the ray counters, in one stream-ordered scratch block
    const size_t table = (size_t)args->heap_size * sizeof(void *);
    std::vector<const void *> host(2 * args->heap_size);
    for (uint32_t i = 0; i < args->heap_size; i++) {
        host[i] = as<BufferObj>(args->vertex_heap[i].id)->ptr;
        host[args->heap_size + i] = as<BufferObj>(args->index_heap[i].id)->ptr;
    }
    uint8_t *scratch = nullptr;
    CUDA_CHECK(cudaMallocAsync((void **)&scratch, 2 * table + 16, s->stream));
    CUDA_CHECK(cudaMemcpyAsync(scratch, host.data(), 2 * table, cudaMemcpyHostToDevice, s->stream));  // pageable source: staged before returning
    CUDA_CHECK(cudaMemsetAsync(scratch + 2 * table, 0, 16, s->stream));
    launch_path_tracer(s->stream, view_of(a), (const float *const *)scratch, (const uint32_t *const *)(scratch + table), (float4 *)img->ptr,
                       (uint32_t *)seed->ptr, args->width, args->height, args->spp_per_dispatch, args->max_depth, args->tan_half_fov,
                       (unsigned long long *)(scratch + 2 * table), d->lc);
    if (ray_counts_out) {
        unsigned long long h[2] = {0, 0};
        CUDA_CHECK(cudaMemcpyAsync(h, scratch + 2 * table, 16, cudaMemcpyDeviceToHost, s->stream));
        CUDA_CHECK(cudaStreamSynchronize(s->stream));
        ray_counts_out[0] = h[0]; ray_counts_out[1] = h[1];
    }
    CUDA_CHECK(cudaFreeAsync(scratch, s->stream));
    flush_launches(d);
}

void lc_b200_trace_closest_counted(lcb_device dev, lcb_stream sh, lcb_accel ah, lcb_buffer rays, size_t rays_offset, lcb_buffer hits, size_t hits_offset,
                                   uint64_t count, uint32_t mask, lcb_trace_counters *out) {
    DeviceObj *d = dev_of(dev); bind(d);
    StreamObj *s = as<StreamObj>(sh.id); AccelObj *a = as<AccelObj>(ah.id);
    BufferObj *rb = as<BufferObj>(rays.id), *hb = as<BufferObj>(hits.id);
    if (rays_offset + count * 32 > rb->size || hits_offset + count * 24 > hb->size) fatal("trace_closest_counted: rays exceed the buffers");
    TraceCounters *ctr = nullptr;
    CUDA_CHECK(cudaMallocAsync((void **)&ctr, sizeof(TraceCounters), s->stream));
    CUDA_CHECK(cudaMemsetAsync(ctr, 0, sizeof(TraceCounters), s->stream));
    if (count) trace_closest(s->stream, view_of(a), rb->ptr + rays_offset, hb->ptr + hits_offset, count, mask, s->work_counter, ctr, d->lc);
    TraceCounters h{};
    CUDA_CHECK(cudaMemcpyAsync(&h, ctr, sizeof(h), cudaMemcpyDeviceToHost, s->stream));
    CUDA_CHECK(cudaFreeAsync(ctr, s->stream));
    CUDA_CHECK(cudaStreamSynchronize(s->stream));
    out->nodes_visited = h.nodes_visited; out->tris_tested = h.tris_tested; out->rays = h.rays; out->instance_entries = h.instance_entries;
    flush_launches(d);
}

// Host-buffer entry points.  The batch is cut into chunks that flow through three lanes — H2D copy, traversal, D2H
// copy — chained by events, so PCIe transfers in both directions overlap the kernels.  Consecutive chunks are traced on
// two alternating streams (each with its own ray-pool counter) so that the tail of one persistent launch overlaps the
// start of the next.
static uint64_t host_chunk_rays() {
    static uint64_t c = [] { uint64_t v = 1ull << 18; if (const char *e = getenv("LC_B200_HOST_CHUNK")) v = strtoull(e, nullptr, 10); return v < 1024 ? 1024 : v; }();
    return c;
}

static uint64_t host_chunk_max_rays() {
    static uint64_t c = [] { uint64_t v = 1ull << 20; if (const char *e = getenv("LC_B200_HOST_CHUNK_MAX")) v = strtoull(e, nullptr, 10); return v; }();
    return c;
}

static unsigned host_grid_limit() {  // LC_B200_HOST_GRID: CTAs per chunk launch of the host pipeline (0 = all resident slots)
    static unsigned c = [] { const char *e = getenv("LC_B200_HOST_GRID"); return e ? (unsigned)strtoul(e, nullptr, 10) : 0u; }();
    return c;
}

static void trace_host_pipelined(DeviceObj *d, AccelObj *a, const lcb_ray *rays, void *out, size_t out_stride, uint64_t count, uint32_t mask, bool any) {
    std::lock_guard<std::mutex> lk(d->mu);
    ensure_stage(d->stage_rays, d->stage_rays_cap, count * 32);
    ensure_stage(d->stage_out, d->stage_out_cap, count * out_stride);
    // Chunk schedule: small chunks at both ends (the first kernel cannot start before its rays arrived, the last download cannot start
    // before its kernel ended), doubling towards the middle where large chunks keep the number of persistent launches down.
    // LC_B200_HOST_CHUNK = size of the end chunks, LC_B200_HOST_CHUNK_MAX = cap in the middle (equal values: uniform chunks).
    const uint64_t c_min = host_chunk_rays(), c_max = std::max(c_min, host_chunk_max_rays());
    std::vector<uint64_t> sizes;
    {
        std::vector<uint64_t> head, tail;
        uint64_t left = count, ch = c_min, ct = c_min;
        while (left > 0) {
            uint64_t n = std::min(ch, left); head.push_back(n); left -= n; ch = std::min(ch * 2, c_max);
            if (left == 0) break;
            n = std::min(ct, left); tail.push_back(n); left -= n; ct = std::min(ct * 2, c_max);
        }
        sizes = head;
        sizes.insert(sizes.end(), tail.rbegin(), tail.rend());
    }
    const size_t n_chunks = sizes.size();
    // timing probes only: 1 no H2D, 2 no D2H, 3 neither — applied from the third call on, so the staged rays are real
    static const int probe_env = [] { const char *e = getenv("LC_B200_HOST_PROBE"); return e ? atoi(e) : 0; }();
    static int probe_calls = 0;
    const int probe = ++probe_calls > 2 ? probe_env : 0;
    std::vector<cudaEvent_t> ev(2 * n_chunks);
    for (auto &e : ev) CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    StreamObj *lanes[2] = {d->internal, d->internal2};
    const AccelView view = view_of(a);
    uint64_t first = 0;
    for (size_t c = 0; c < n_chunks; c++) {
        const uint64_t n = sizes[c];
        StreamObj *k = lanes[c & 1];
        if (!(probe & 1)) CUDA_CHECK(cudaMemcpyAsync(d->stage_rays + first * 32, rays + first, n * 32, cudaMemcpyHostToDevice, d->copy_in));
        CUDA_CHECK(cudaEventRecord(ev[2 * c], d->copy_in));
        CUDA_CHECK(cudaStreamWaitEvent(k->stream, ev[2 * c], 0));
        if (any) trace_any(k->stream, view, d->stage_rays + first * 32, (uint32_t *)(d->stage_out + first * out_stride), n, mask, k->work_counter, d->lc, host_grid_limit());
        else trace_closest(k->stream, view, d->stage_rays + first * 32, d->stage_out + first * out_stride, n, mask, k->work_counter, nullptr, d->lc, host_grid_limit());
        CUDA_CHECK(cudaEventRecord(ev[2 * c + 1], k->stream));
        CUDA_CHECK(cudaStreamWaitEvent(d->copy_out, ev[2 * c + 1], 0));
        if (!(probe & 2)) CUDA_CHECK(cudaMemcpyAsync((uint8_t *)out + first * out_stride, d->stage_out + first * out_stride, n * out_stride, cudaMemcpyDeviceToHost, d->copy_out));
        first += n;
    }
    CUDA_CHECK(cudaStreamSynchronize(d->copy_out));
    CUDA_CHECK(cudaStreamSynchronize(lanes[0]->stream));
    CUDA_CHECK(cudaStreamSynchronize(lanes[1]->stream));
    for (auto &e : ev) cudaEventDestroy(e);
    flush_launches(d);
}

void lc_b200_trace_closest_host(lcb_device dev, lcb_accel ah, const lcb_ray *rays, lcb_surface_hit *hits, uint64_t count, uint32_t mask) {
    DeviceObj *d = dev_of(dev); bind(d);
    if (!count) return;
    trace_host_pipelined(d, as<AccelObj>(ah.id), rays, hits, 24, count, mask, false);
}

void lc_b200_trace_any_host(lcb_device dev, lcb_accel ah, const lcb_ray *rays, uint32_t *occluded, uint64_t count, uint32_t mask) {
    DeviceObj *d = dev_of(dev); bind(d);
    if (!count) return;
    trace_host_pipelined(d, as<AccelObj>(ah.id), rays, occluded, 4, count, mask, true);
}

void lc_b200_instance_transform(lcb_device, lcb_accel ah, uint32_t i, float *out) {
    AccelObj *a = as<AccelObj>(ah.id); std::lock_guard<std::mutex> lk(a->mu);
    if (i >= a->instances.size()) fatal("instance_transform: index %u out of range", i);
    memcpy(out, a->instances[i].affine, 48);
}
uint32_t lc_b200_instance_user_id(lcb_device, lcb_accel ah, uint32_t i) {
    AccelObj *a = as<AccelObj>(ah.id); std::lock_guard<std::mutex> lk(a->mu);
    if (i >= a->instances.size()) fatal("instance_user_id: index %u out of range", i);
    return a->instances[i].user_id;
}
uint32_t lc_b200_instance_visibility_mask(lcb_device, lcb_accel ah, uint32_t i) {
    AccelObj *a = as<AccelObj>(ah.id); std::lock_guard<std::mutex> lk(a->mu);
    if (i >= a->instances.size()) fatal("instance_visibility_mask: index %u out of range", i);
    if (!a->instances[i].valid) fatal("instance_visibility_mask: empty instance slot %u", i);  // accel.rs:556
    return a->instances[i].visible;
}

void lc_b200_mesh_stats(lcb_device dev, lcb_mesh h, lcb_build_stats *out) { bind(dev_of(dev)); MeshObj *m = as<MeshObj>(h.id); std::lock_guard<std::mutex> lk(m->mu); mesh_finalize(m); *out = m->stats; }
void lc_b200_accel_stats(lcb_device dev, lcb_accel h, lcb_build_stats *out) { bind(dev_of(dev)); AccelObj *a = as<AccelObj>(h.id); std::lock_guard<std::mutex> lk(a->mu); accel_finalize(a); *out = a->stats; }

void *lc_b200_stream_native(lcb_device, lcb_stream h) { return (void *)as<StreamObj>(h.id)->stream; }
void *lc_b200_buffer_native(lcb_device, lcb_buffer h) { return as<BufferObj>(h.id)->ptr; }
int lc_b200_device_ordinal(lcb_device dev) { return dev_of(dev)->ordinal; }
uint64_t lc_b200_kernel_launch_count(void) { return g_launches.load(); }
const char *lc_b200_version(void) { return "lc_b200 0.2 (sm_100a; LBVH / PLOC + 8-wide quantised BVH; canonical fp32 watertight traversal; IR -> CUDA lowering via NVRTC)"; }

// The CUDA source the lowering produces for a KernelModule (malloc'd; release with free_string).  No GPU needed.
char *lc_b200_ir_lower_source(const void *kernel_module) {
    LoweredKernel k;
    try { lower_kernel(reinterpret_cast<const ir::KernelModule *>(kernel_module), k); } catch (const std::exception &e) { fatal("%s", e.what()); }
    char *out = (char *)malloc(k.source.size() + 1); memcpy(out, k.source.c_str(), k.source.size() + 1); return out;
}
// Lower + NVRTC-compile without loading the module (no GPU needed).  Returns 0 on success; *log (malloc'd, may be empty) holds the
// lowering diagnostic or the NVRTC log.
int lc_b200_shader_compile_check(const void *kernel_module, bool fast_math, char **log) {
    std::string l; int rc = 0;
    try { ShaderObj *s = shader_create(reinterpret_cast<const ir::KernelModule *>(kernel_module), fast_math, true, nullptr, l); shader_destroy(s); }
    catch (const std::exception &e) { l = e.what(); rc = 1; }
    if (log) { *log = (char *)malloc(l.size() + 1); memcpy(*log, l.c_str(), l.size() + 1); }
    return rc;
}

int lc_b200_set_builder(int builder) {
    if (builder < -1 || builder > kBuilderAuto) fatal("lc_b200_set_builder: unknown builder %d", builder);
    return g_builder_override.exchange(builder);
}

const void *lc_b200_make_ir_type(size_t size, size_t alignment) {
    // leaked on purpose: type blocks live for the process, like the frontend's interned types
    auto *ty = new IrType; memset(ty, 0, sizeof(*ty));
    if (size == 0) ty->tag = IR_VOID;
    else { ty->tag = IR_STRUCT; ty->u.struct_.fields = IrSlice{nullptr, 0, nullptr}; ty->u.struct_.alignment = alignment; ty->u.struct_.size = size; }
    auto *blk = new IrArcBlock; blk->ptr = ty; blk->ref_count.store(1); blk->destructor = nullptr;
    auto **arc = new IrArcBlock *; *arc = blk;
    return arc;
}

}  // extern "C"
