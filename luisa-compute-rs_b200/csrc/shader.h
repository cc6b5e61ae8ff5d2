// shader.h — internal interface of the IR -> CUDA lowering (ir_lower.cpp) and the NVRTC-backed shader objects (shader.cu).
// Row 9 of SURVEY.md §8a: how DSL kernel bodies reach the traversal routines.  The reference CPU backend does the same job in
// luisa_compute_backend_impl/src/cpu/codegen/cpp.rs (IR -> C++ text) + cpu/shader.rs (clang) + cpu/stream.rs:330-440 (launch).
#pragma once
#include "ir_layout.h"
#include "rt_types.cuh"

#include <string>
#include <vector>

namespace lcb {

// One kernel parameter slot: a capture (binding fixed at create_shader time) or an argument (bound per ShaderDispatch).
struct ParamSlot {
    enum Kind { Buffer, Texture, Bindless, Accel, Uniform } kind;
    size_t offset = 0;        // byte offset inside the kernel parameter block (lc_params)
    size_t size = 0;          // Uniform: byte size of the value
    bool is_capture = false;
    ir::Binding binding{};    // captures only
};

struct LoweredKernel {
    std::string source;               // CUDA C++ translation unit (includes lc_device_lib.cuh)
    std::vector<ParamSlot> captures;  // in KernelModule.captures order
    std::vector<ParamSlot> args;      // in KernelModule.args order
    size_t param_bytes = 0;           // sizeof(lc_params)
    uint32_t block_size[3] = {1, 1, 1};
    bool wave = false;                // wavefront lowering: persistent grid of kWaveThreads-thread CTAs pulling dispatch ids from a work counter
    bool writes_accel = false;        // contains RayTracingSetInstance*: AccelBuild must read the instance table back before it applies modifications
    int wave_yield_min = 8;           // ready lanes per warp at which the traversal loop hands control back to the kernel body
    std::vector<std::string> messages;  // assert / unreachable texts, indexed by the id the kernel prints
};

// Throws std::runtime_error with a diagnostic naming the unsupported construct.
void lower_kernel(const ir::KernelModule *km, LoweredKernel &out);

// Host images of the device-side parameter records (lc_device_lib.cuh: lc_buffer, lc_texture, lc_bindless, lc_accel, lc_launch)
struct HostBufferArg { void *ptr; uint64_t size; };
struct HostTextureArg { void *data; uint32_t width, height, depth; uint32_t storage; uint32_t sampler; uint32_t pad; };  // sampler = filter | address << 2 (Sampler::encode, api_types:448-452)
struct HostBindlessSlot { void *buffer; uint64_t buffer_size; HostTextureArg tex2d, tex3d; };
struct HostBindlessArg { const HostBindlessSlot *slots; uint64_t count; };
struct HostAccelArg { AccelView view; InstanceRec *instances_rw; uint32_t *dirty; };  // dirty: set by kernels that edit the instance table
// yield_min / work_counter / work_items: wavefront-lowered kernels only (ir_lower.cpp header; trace_device.cuh wave_fetch / wave_traverse)
struct HostLaunch { uint32_t dispatch_size[3]; uint32_t yield_min; unsigned long long *work_counter; unsigned long long work_items; };
static_assert(sizeof(HostBufferArg) == 16 && sizeof(HostTextureArg) == 32 && sizeof(HostBindlessSlot) == 80 && sizeof(HostBindlessArg) == 16 &&
                  sizeof(HostAccelArg) == 72 && sizeof(HostLaunch) == 32,
              "parameter records are mirrored byte for byte in lc_device_lib.cuh");

// NVRTC + module loading (shader.cu).  compile_only: stop after NVRTC (usable without a GPU; create_shader's compile_only option).
struct ShaderObj;
ShaderObj *shader_create(const ir::KernelModule *km, bool fast_math, bool compile_only, const char *name, std::string &log);
void shader_destroy(ShaderObj *);
const LoweredKernel &shader_lowered(const ShaderObj *);
// params: a filled lc_params image of shader_lowered().param_bytes bytes
// work_counter: an 8-byte device word owned by the stream (zeroed in stream order before a wavefront-lowered kernel starts).
// Returns the number of kernels launched.
int shader_launch(ShaderObj *, cudaStream_t stream, void *params, const uint32_t dispatch_size[3], unsigned long long *work_counter);

}  // namespace lcb
