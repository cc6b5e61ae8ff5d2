// liblc-backend-b200.so — the C++-ABI face of the B200 device.
//
// C++ LuisaCompute loads a backend as `lc-backend-<name>` (on Linux `liblc-backend-<name>.so`, LC/src/core/platform.cpp) from the runtime directory and binds three symbols,
// `create`, `destroy`, `backend_device_names` (LC/src/runtime/context.cpp:72-92; signatures device.h:104-105).
// What comes back is a `luisa::compute::DeviceInterface` (LC/include/luisa/runtime/rhi/device_interface.h:99-220)
// whose virtuals take C++ objects (`CommandList`, `Function`, `const Type *`).  This file is the adapter from those
// virtuals to the C function-pointer table of include/lc_b200_api.h — the role LC's own rust_device_common.cpp
// (`src/backends/common/`) plays for the Rust CPU backend.  No device logic lives here: every call lands in
// liblc_b200.so, which this library links with an $ORIGIN rpath.
//
// It is compiled against the reference's own headers where they lie (make -C csrc cpp_backend REF=...; development
// container only — the headers are not copied into this repository) and resolves `luisa::detail::allocator_*`,
// `luisa::log_*`, `AST2IR::*`, `DeviceInterface::DeviceInterface` from the host program's lc-core / lc-runtime / lc-ir
// at load time, exactly like every other LuisaCompute backend module.
#include <cstring>
#include <limits>

#include <luisa/core/logging.h>
#include <luisa/core/stl/vector.h>
#include <luisa/core/stl/string.h>
#include <luisa/runtime/context.h>
#include <luisa/runtime/rhi/device_interface.h>
#include <luisa/runtime/rhi/command.h>
#include <luisa/runtime/rhi/sampler.h>
#include <luisa/runtime/command_list.h>
#include <luisa/ir/ast2ir.h>

#include "../../include/lc_b200_api.h"

namespace lc_b200_cpp {

using namespace luisa;
using namespace luisa::compute;

// The C++ enums and the C ABI's come from the same definitions (api_types:285-407 are generated from pixel.h's order).
static_assert((int)PixelStorage::FLOAT4 == 14 && (int)PixelStorage::BC7 == 23, "PixelStorage order");
static_assert((int)PixelFormat::RGBA32F == 29 && (int)PixelFormat::RGBA8UNorm == 8, "PixelFormat order");
static_assert((int)AccelBuildRequest::FORCE_BUILD == LCB_REQUEST_FORCE_BUILD, "AccelBuildRequest order");
static_assert((int)AccelOption::UsageHint::FAST_BUILD == LCB_HINT_FAST_BUILD, "AccelUsageHint order");
static_assert((int)StreamTag::GRAPHICS == LCB_STREAM_GRAPHICS && (int)StreamTag::COMPUTE == LCB_STREAM_COMPUTE &&
                  (int)StreamTag::COPY == LCB_STREAM_COPY, "StreamTag order");
static_assert(AccelBuildCommand::Modification::flag_primitive == LCB_MOD_PRIMITIVE &&
                  AccelBuildCommand::Modification::flag_transform == LCB_MOD_TRANSFORM &&
                  AccelBuildCommand::Modification::flag_opaque_on == LCB_MOD_OPAQUE_ON &&
                  AccelBuildCommand::Modification::flag_opaque_off == LCB_MOD_OPAQUE_OFF &&
                  AccelBuildCommand::Modification::flag_visibility == LCB_MOD_VISIBILITY &&
                  AccelBuildCommand::Modification::flag_user_id == LCB_MOD_USER_ID, "modification flag bits");
static_assert((int)Argument::Tag::BUFFER == LCB_ARG_BUFFER && (int)Argument::Tag::TEXTURE == LCB_ARG_TEXTURE &&
                  (int)Argument::Tag::UNIFORM == LCB_ARG_UNIFORM && (int)Argument::Tag::BINDLESS_ARRAY == LCB_ARG_BINDLESS &&
                  (int)Argument::Tag::ACCEL == LCB_ARG_ACCEL, "Argument tag order");

static lcb_accel_option to_c(const AccelOption &o) noexcept {
    return lcb_accel_option{(int32_t)o.hint, o.allow_compaction, o.allow_update};
}
static ResourceCreationInfo to_cpp(lcb_created c) noexcept {
    ResourceCreationInfo info{};
    info.handle = c.handle;
    info.native_handle = c.native_handle;
    return info;
}

// One submitted CommandList.  The C table consumes the lcb_command array synchronously but reads the nested arrays
// (arguments, uniform bytes, modifications, download targets) until the completion callback (SURVEY.md §8b
// "Ownership"), so the C++ commands — which own that storage — and the converted side arrays live here until then.
struct Submission final : public CommandVisitor {
    CommandList::CommandContainer owned;
    CommandList::CallbackContainer callbacks;
    luisa::vector<lcb_command> out;
    luisa::vector<luisa::vector<lcb_argument>> arg_arrays;
    luisa::vector<luisa::vector<lcb_accel_modification>> accel_mods;
    luisa::vector<luisa::vector<lcb_bindless_modification>> bindless_mods;

    static void require_no_offset(uint3 off, const char *what) noexcept {
        // api::Texture*Command carries no region origin (api_types:520-575); the C++ API can express one.
        if (off.x | off.y | off.z) { LUISA_ERROR_WITH_LOCATION("b200: {} with a non-zero texture offset is not expressible through the device ABI.", what); }
    }
    lcb_command &push(int32_t tag) noexcept {
        lcb_command c;
        std::memset(&c, 0, sizeof(c));
        c.tag = tag;
        out.push_back(c);
        return out.back();
    }

    void visit(const BufferUploadCommand *c) noexcept override {
        push(LCB_CMD_BUFFER_UPLOAD).u.buffer_upload = {{c->handle()}, c->offset(), c->size(), static_cast<const uint8_t *>(c->data())};
    }
    void visit(const BufferDownloadCommand *c) noexcept override {
        push(LCB_CMD_BUFFER_DOWNLOAD).u.buffer_download = {{c->handle()}, c->offset(), c->size(), static_cast<uint8_t *>(c->data())};
    }
    void visit(const BufferCopyCommand *c) noexcept override {
        push(LCB_CMD_BUFFER_COPY).u.buffer_copy = {{c->src_handle()}, c->src_offset(), {c->dst_handle()}, c->dst_offset(), c->size()};
    }
    void visit(const BufferToTextureCopyCommand *c) noexcept override {
        require_no_offset(c->texture_offset(), "BufferToTextureCopyCommand");
        auto s = c->size();
        push(LCB_CMD_BUFFER_TO_TEXTURE).u.buffer_to_texture = {{c->buffer()}, c->buffer_offset(), {c->texture()}, (int32_t)c->storage(), c->level(), {s.x, s.y, s.z}};
    }
    void visit(const TextureToBufferCopyCommand *c) noexcept override {
        require_no_offset(c->texture_offset(), "TextureToBufferCopyCommand");
        auto s = c->size();
        push(LCB_CMD_TEXTURE_TO_BUFFER).u.texture_to_buffer = {{c->buffer()}, c->buffer_offset(), {c->texture()}, (int32_t)c->storage(), c->level(), {s.x, s.y, s.z}};
    }
    void visit(const TextureUploadCommand *c) noexcept override {
        require_no_offset(c->offset(), "TextureUploadCommand");
        auto s = c->size();
        push(LCB_CMD_TEXTURE_UPLOAD).u.texture_upload = {{c->handle()}, (int32_t)c->storage(), c->level(), {s.x, s.y, s.z},
                                                          const_cast<uint8_t *>(static_cast<const uint8_t *>(c->data()))};
    }
    void visit(const TextureDownloadCommand *c) noexcept override {
        require_no_offset(c->offset(), "TextureDownloadCommand");
        auto s = c->size();
        push(LCB_CMD_TEXTURE_DOWNLOAD).u.texture_download = {{c->handle()}, (int32_t)c->storage(), c->level(), {s.x, s.y, s.z}, static_cast<uint8_t *>(c->data())};
    }
    void visit(const TextureCopyCommand *c) noexcept override {
        require_no_offset(c->src_offset(), "TextureCopyCommand (source)");
        require_no_offset(c->dst_offset(), "TextureCopyCommand (destination)");
        auto s = c->size();
        push(LCB_CMD_TEXTURE_COPY).u.texture_copy = {(int32_t)c->storage(), {c->src_handle()}, {c->dst_handle()}, {s.x, s.y, s.z}, c->src_level(), c->dst_level()};
    }
    void visit(const ShaderDispatchCommand *c) noexcept override {
        if (c->is_indirect() || c->is_multiple_dispatch()) {
            LUISA_ERROR_WITH_LOCATION("b200: indirect / multi-size dispatch has no counterpart in the device ABI (api_types:586-601).");
        }
        auto &args = arg_arrays.emplace_back();
        args.reserve(c->arguments().size());
        for (auto &&a : c->arguments()) {
            lcb_argument x;
            std::memset(&x, 0, sizeof(x));
            x.tag = (int32_t)a.tag;
            switch (a.tag) {
                case Argument::Tag::BUFFER: x.u.buffer = {{a.buffer.handle}, a.buffer.offset, a.buffer.size}; break;
                case Argument::Tag::TEXTURE: x.u.texture = {{a.texture.handle}, a.texture.level}; break;
                case Argument::Tag::UNIFORM: {
                    // uniforms are (offset, size) into the command's own argument buffer: hand out the bytes in place
                    auto bytes = c->uniform(a.uniform);
                    x.u.uniform = {reinterpret_cast<const uint8_t *>(bytes.data()), bytes.size()};
                    break;
                }
                case Argument::Tag::BINDLESS_ARRAY: x.u.bindless = {a.bindless_array.handle}; break;
                case Argument::Tag::ACCEL: x.u.accel = {a.accel.handle}; break;
            }
            args.push_back(x);
        }
        auto n = c->dispatch_size();
        push(LCB_CMD_SHADER_DISPATCH).u.shader_dispatch = {{c->handle()}, {n.x, n.y, n.z}, args.data(), args.size()};
    }
    void visit(const MeshBuildCommand *c) noexcept override {
        auto &m = push(LCB_CMD_MESH_BUILD).u.mesh_build;
        m.mesh = {c->handle()};
        m.request = (int32_t)c->request();
        m.vertex_buffer = {c->vertex_buffer()};
        m.vertex_buffer_offset = c->vertex_buffer_offset();
        m.vertex_buffer_size = c->vertex_buffer_size();
        m.vertex_stride = c->vertex_stride();
        m.index_buffer = {c->triangle_buffer()};
        m.index_buffer_offset = c->triangle_buffer_offset();
        m.index_buffer_size = c->triangle_buffer_size();
        m.index_stride = 3 * sizeof(uint32_t);// the only stride the runtime accepts (LC/src/api/runtime.cpp:191)
    }
    void visit(const CurveBuildCommand *c) noexcept override {
        push(LCB_CMD_CURVE_BUILD).u.curve_build = {{c->handle()}, (int32_t)c->request(), (int32_t)c->basis(), c->cp_count(), c->seg_count(),
                                                    {c->cp_buffer()}, c->cp_buffer_offset(), c->cp_stride(), {c->seg_buffer()}, c->seg_buffer_offset()};
    }
    void visit(const ProceduralPrimitiveBuildCommand *c) noexcept override {
        // the C++ command states the AABB range in bytes, the device ABI in boxes (api_types:633-641)
        push(LCB_CMD_PROCEDURAL_BUILD).u.procedural_build = {{c->handle()}, (int32_t)c->request(), {c->aabb_buffer()},
                                                              c->aabb_buffer_offset(), c->aabb_buffer_size() / sizeof(lcb_aabb)};
    }
    void visit(const AccelBuildCommand *c) noexcept override {
        auto &mods = accel_mods.emplace_back();
        mods.reserve(c->modifications().size());
        for (auto &&m : c->modifications()) {
            lcb_accel_modification x;
            x.index = m.index;
            x.user_id = m.user_id;
            x.flags = m.flags;
            x.visibility = m.vis_mask;
            x.mesh = m.primitive;
            std::memcpy(x.affine, m.affine, sizeof(x.affine));
            mods.push_back(x);
        }
        auto &a = push(LCB_CMD_ACCEL_BUILD).u.accel_build;
        a.accel = {c->handle()};
        a.request = (int32_t)c->request();
        a.instance_count = c->instance_count();
        a.modifications = mods.data();
        a.modifications_count = mods.size();
        a.update_instance_buffer_only = c->update_instance_buffer_only();
    }
    void visit(const BindlessArrayUpdateCommand *c) noexcept override {
        using Mod = BindlessArrayUpdateCommand::Modification;
        auto tex = [](const Mod::Texture &t) noexcept {
            return lcb_bindless_texture_update{(int32_t)t.op, {t.handle}, {(int32_t)t.sampler.filter(), (int32_t)t.sampler.address()}};
        };
        auto &mods = bindless_mods.emplace_back();
        mods.reserve(c->modifications().size());
        for (auto &&m : c->modifications()) {
            mods.push_back(lcb_bindless_modification{m.slot, {(int32_t)m.buffer.op, {m.buffer.handle}, m.buffer.offset_bytes}, tex(m.tex2d), tex(m.tex3d)});
        }
        push(LCB_CMD_BINDLESS_UPDATE).u.bindless_update = {{c->handle()}, mods.data(), mods.size()};
    }
    void visit(const CustomCommand *c) noexcept override {
        LUISA_ERROR_WITH_LOCATION("b200: custom command {:#x} is not supported.", c->uuid());
    }

    // completion: runs on the stream's host thread, exactly once (cpu/stream.rs:133-141)
    static void completed(uint8_t *ctx) noexcept {
        auto self = reinterpret_cast<Submission *>(ctx);
        for (auto &&f : self->callbacks) { f(); }
        luisa::delete_with_allocator(self);
    }
};

class B200Device final : public DeviceInterface {
    lcb_lib_interface _lib;
    lcb_context _lib_ctx;
    lcb_device_interface _d;

public:
    B200Device(Context &&ctx, const DeviceConfig *config) noexcept : DeviceInterface{std::move(ctx)} {
        _lib = luisa_compute_lib_interface();
        _lib.set_logger_callback([](lcb_logger_message m) {
            // levels are the `log` crate's initials (backend_impl/src/lib.rs:101-131)
            switch (m.level ? m.level[0] : 'I') {
                case 'E': LUISA_ERROR("[{}] {}", m.target, m.message); break;
                case 'W': LUISA_WARNING("[{}] {}", m.target, m.message); break;
                case 'D': case 'T': LUISA_VERBOSE("[{}] {}", m.target, m.message); break;
                default: LUISA_INFO("[{}] {}", m.target, m.message); break;
            }
        });
        auto dir = luisa::to_string(context().runtime_directory());
        _lib_ctx = _lib.create_context(dir.c_str());
        auto json = config && config->device_index != std::numeric_limits<size_t>::max() ?
                        luisa::format("{{\"device_index\": {}}}", config->device_index) :
                        luisa::string{"{}"};
        _d = _lib.create_device(_lib_ctx, "b200", json.c_str());
    }
    ~B200Device() noexcept override {
        _d.destroy_device(_d);
        _lib.destroy_context(_lib_ctx);
    }

    void *native_handle() const noexcept override { return _d.native_handle(_d.device); }
    uint compute_warp_size() const noexcept override { return _d.compute_warp_size(_d.device); }

    BufferCreationInfo create_buffer(const Type *element, size_t elem_count, void *external_memory) noexcept override {
        auto ir_type = AST2IR::build_type(element);
        return create_buffer(&ir_type, elem_count, external_memory);
    }
    BufferCreationInfo create_buffer(const ir::CArc<ir::Type> *element, size_t elem_count, void *external_memory) noexcept override {
        auto c = _d.create_buffer(_d.device, element, elem_count, external_memory);
        BufferCreationInfo info{};
        info.handle = c.resource.handle;
        info.native_handle = c.resource.native_handle;
        info.element_stride = c.element_stride;
        info.total_size_bytes = c.total_size_bytes;
        return info;
    }
    void destroy_buffer(uint64_t handle) noexcept override { _d.destroy_buffer(_d.device, {handle}); }

    ResourceCreationInfo create_texture(PixelFormat format, uint dimension, uint width, uint height, uint depth,
                                        uint mipmap_levels, bool simultaneous_access, bool allow_raster_target) noexcept override {
        return to_cpp(_d.create_texture(_d.device, (int32_t)format, dimension, width, height, depth, mipmap_levels, simultaneous_access, allow_raster_target));
    }
    void destroy_texture(uint64_t handle) noexcept override { _d.destroy_texture(_d.device, {handle}); }

    ResourceCreationInfo create_bindless_array(size_t size) noexcept override { return to_cpp(_d.create_bindless_array(_d.device, size)); }
    void destroy_bindless_array(uint64_t handle) noexcept override { _d.destroy_bindless_array(_d.device, {handle}); }

    ResourceCreationInfo create_stream(StreamTag tag) noexcept override { return to_cpp(_d.create_stream(_d.device, (int32_t)tag)); }
    void destroy_stream(uint64_t handle) noexcept override { _d.destroy_stream(_d.device, {handle}); }
    void synchronize_stream(uint64_t handle) noexcept override { _d.synchronize_stream(_d.device, {handle}); }

    void dispatch(uint64_t stream, CommandList &&list) noexcept override {
        auto s = luisa::new_with_allocator<Submission>();
        s->owned = list.steal_commands();
        s->callbacks = list.steal_callbacks();
        s->out.reserve(s->owned.size());
        // side arrays must not move once a command points into them
        s->arg_arrays.reserve(s->owned.size());
        s->accel_mods.reserve(s->owned.size());
        s->bindless_mods.reserve(s->owned.size());
        for (auto &&c : s->owned) { c->accept(*s); }
        _d.dispatch(_d.device, {stream}, lcb_command_list{s->out.data(), s->out.size()}, &Submission::completed, reinterpret_cast<uint8_t *>(s));
    }

    SwapchainCreationInfo create_swapchain(const SwapchainOption &, uint64_t) noexcept override {
        LUISA_ERROR_WITH_LOCATION("b200: swapchains are outside the device's scope (DESIGN.md §7).");
    }
    void destroy_swap_chain(uint64_t) noexcept override {}
    void present_display_in_stream(uint64_t, uint64_t, uint64_t) noexcept override {
        LUISA_ERROR_WITH_LOCATION("b200: swapchains are outside the device's scope (DESIGN.md §7).");
    }

    ShaderCreationInfo create_shader(const ShaderOption &option, Function kernel) noexcept override {
        // the device consumes the SSA IR (csrc/ir_lower.cpp); the AST is converted by LuisaCompute's own AST2IR
        auto module = AST2IR::build_kernel(kernel);
        return create_shader(option, module->get());
    }
    ShaderCreationInfo create_shader(const ShaderOption &option, const ir::KernelModule *kernel) noexcept override {
        lcb_shader_option o{};
        o.enable_cache = option.enable_cache;
        o.enable_fast_math = option.enable_fast_math;
        o.enable_debug_info = option.enable_debug_info;
        o.compile_only = option.compile_only;
        o.time_trace = option.time_trace;
        o.max_registers = option.max_registers;
        o.name = option.name.c_str();
        o.native_include = option.native_include.c_str();
        auto c = _d.create_shader(_d.device, lcb_kernel_module{reinterpret_cast<uint64_t>(kernel)}, &o);
        ShaderCreationInfo info{};
        info.handle = c.resource.handle;
        info.native_handle = c.resource.native_handle;
        info.block_size = make_uint3(c.block_size[0], c.block_size[1], c.block_size[2]);
        return info;
    }
    ShaderCreationInfo load_shader(luisa::string_view name, luisa::span<const Type *const>) noexcept override {
        LUISA_ERROR_WITH_LOCATION("b200: no ahead-of-time shader store; '{}' must be created with create_shader.", name);
    }
    Usage shader_argument_usage(uint64_t, size_t) noexcept override { return Usage::READ_WRITE; }
    void destroy_shader(uint64_t handle) noexcept override { _d.destroy_shader(_d.device, {handle}); }

    ResourceCreationInfo create_event() noexcept override { return to_cpp(_d.create_event(_d.device)); }
    void destroy_event(uint64_t handle) noexcept override { _d.destroy_event(_d.device, {handle}); }
    void signal_event(uint64_t handle, uint64_t stream, uint64_t value) noexcept override { _d.signal_event(_d.device, {handle}, {stream}, value); }
    void wait_event(uint64_t handle, uint64_t stream, uint64_t value) noexcept override { _d.wait_event(_d.device, {handle}, {stream}, value); }
    bool is_event_completed(uint64_t handle, uint64_t value) const noexcept override { return _d.is_event_completed(_d.device, {handle}, value); }
    void synchronize_event(uint64_t handle, uint64_t value) noexcept override { _d.synchronize_event(_d.device, {handle}, value); }

    ResourceCreationInfo create_mesh(const AccelOption &option) noexcept override {
        auto o = to_c(option);
        return to_cpp(_d.create_mesh(_d.device, &o));
    }
    void destroy_mesh(uint64_t handle) noexcept override { _d.destroy_mesh(_d.device, {handle}); }
    ResourceCreationInfo create_procedural_primitive(const AccelOption &option) noexcept override {
        auto o = to_c(option);
        return to_cpp(_d.create_procedural_primitive(_d.device, &o));
    }
    void destroy_procedural_primitive(uint64_t handle) noexcept override { _d.destroy_procedural_primitive(_d.device, {handle}); }
    ResourceCreationInfo create_curve(const AccelOption &option) noexcept override {
        auto o = to_c(option);
        return to_cpp(_d.create_curve(_d.device, &o));
    }
    void destroy_curve(uint64_t handle) noexcept override { _d.destroy_curve(_d.device, {handle}); }
    ResourceCreationInfo create_accel(const AccelOption &option) noexcept override {
        auto o = to_c(option);
        return to_cpp(_d.create_accel(_d.device, &o));
    }
    void destroy_accel(uint64_t handle) noexcept override { _d.destroy_accel(_d.device, {handle}); }

    luisa::string query(luisa::string_view property) noexcept override {
        luisa::string key{property};
        auto s = _d.query(_d.device, key.c_str());
        if (s == nullptr) { return {}; }
        luisa::string r{s};
        _lib.free_string(s);
        return r;
    }
    void set_name(Resource::Tag, uint64_t, luisa::string_view) noexcept override {}
};

}// namespace lc_b200_cpp

LUISA_EXPORT_API luisa::compute::DeviceInterface *create(luisa::compute::Context &&ctx, const luisa::compute::DeviceConfig *config) noexcept {
    return luisa::new_with_allocator<lc_b200_cpp::B200Device>(std::move(ctx), config);
}
LUISA_EXPORT_API void destroy(luisa::compute::DeviceInterface *device) noexcept {
    luisa::delete_with_allocator(static_cast<lc_b200_cpp::B200Device *>(device));
}
LUISA_EXPORT_API void backend_device_names(luisa::vector<luisa::string> &names) noexcept {
    names.clear();
    names.emplace_back("NVIDIA B200 (sm_100a)");
}
