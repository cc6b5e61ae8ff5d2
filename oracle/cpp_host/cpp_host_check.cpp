// Test infrastructure: a C++ LuisaCompute host program for lc-backend-b200.so.
//
// What is the reference's and what is ours: `luisa::compute::Context`, `DynamicModule`, `CommandList`, the `*Command` classes and
// `DeviceInterface` are the reference's own code, compiled from LC/src/{core,runtime} where it lies (oracle/Makefile `cpp_host`);
// the backend module the Context finds and loads from its runtime directory is csrc/cpp_backend.cpp + liblc_b200.so.  The program
// drives the path exactly as LC's C++ runtime would — Context::create_device("b200"), buffers, one CommandList carrying uploads +
// MeshBuildCommand + AccelBuildCommand (+ a host callback), a second list after a PREFER_UPDATE vertex edit and an instance
// modification, then a DSL kernel (create_shader from an ir::KernelModule + ShaderDispatchCommand) — traces a ray grid through the batch entry point, downloads the hits with a BufferDownloadCommand and compares
// them bit for bit with the CPU oracle (liboracle.so) on the same scene.  Exit code 0 and a final "ok" line = pass.
#include <atomic>
#include <cstdio>
#include <cstring>
#include <vector>

#include <luisa/core/logging.h>
#include <luisa/runtime/context.h>
#include <luisa/runtime/device.h>
#include <luisa/runtime/rhi/device_interface.h>
#include <luisa/runtime/rhi/command.h>
#include <luisa/runtime/command_list.h>
#include <luisa/runtime/rhi/command_encoder.h>
#include <luisa/runtime/rhi/sampler.h>
#include <luisa/runtime/rhi/pixel.h>
#include <luisa/rust/ir.hpp>

#include "../ir_ref_build.hpp"

#include "../../include/lc_b200_api.h"
extern "C" {
#include "../oracle.h"
}

using namespace luisa;
using namespace luisa::compute;

static ir::CArc<ir::Type> byte_buffer_type() {// Type::Void = raw bytes (cpu/mod.rs:51-83)
    auto t = new ir::Type{};
    t->tag = ir::Type::Tag::Void;
    return ir::CArc<ir::Type>{new ir::CArcSharedBlock<ir::Type>{t, {1}, nullptr}};
}

#define CHECK(c) do { if (!(c)) { std::fprintf(stderr, "cpp_host_check FAILED at line %d: %s\n", __LINE__, #c); return 1; } } while (0)

int main(int argc, char **argv) {
    log_level_warning();
    Context ctx{argv[0]};
    bool installed = false;
    for (auto &&b : ctx.installed_backends()) { installed |= (b == "b200"); }
    CHECK(installed);
    auto names = ctx.backend_device_names("b200");
    CHECK(names.size() == 1);
    Device device = ctx.create_device("b200");
    DeviceInterface *d = device.impl();
    CHECK(d->backend_name() == "b200");
    CHECK(d->query("device_name") == "b200");
    CHECK(d->compute_warp_size() == 32u);
    std::printf("device: %s / %s\n", names[0].c_str(), luisa::string{d->query("device_name")}.c_str());

    // the batch ray entry points live beside the table in liblc_b200.so (include/lc_b200_api.h)
    auto trace_closest = &lc_b200_trace_closest;
    auto trace_any = &lc_b200_trace_any;
    lcb_device cdev{reinterpret_cast<uint64_t>(d->native_handle())};

    // scene: a 32 x 32 grid of quads (2048 triangles) with a ripple, instanced twice (second one lifted and rotated)
    constexpr uint32_t G = 32;
    std::vector<float> verts;
    std::vector<uint32_t> idx;
    auto fill_vertices = [&](float phase) {
        verts.clear();
        for (uint32_t y = 0; y <= G; y++)
            for (uint32_t x = 0; x <= G; x++) {
                float fx = (float)x / G - 0.5f, fy = (float)y / G - 0.5f;
                verts.insert(verts.end(), {fx, fy, 0.05f * std::sin(9.0f * fx + phase) * std::cos(7.0f * fy)});
            }
    };
    fill_vertices(0.f);
    for (uint32_t y = 0; y < G; y++)
        for (uint32_t x = 0; x < G; x++) {
            uint32_t a = y * (G + 1) + x, b = a + 1, c = a + G + 1, e = c + 1;
            idx.insert(idx.end(), {a, b, e, a, e, c});
        }
    const uint32_t n_tri = (uint32_t)idx.size() / 3;
    constexpr uint32_t W = 96, H = 96, N = W * H;
    std::vector<oracle_ray> rays(N);
    for (uint32_t i = 0; i < N; i++) {
        float px = ((i % W) + 0.37f) / W - 0.5f, py = ((i / W) + 0.61f) / H - 0.5f;
        rays[i] = oracle_ray{{0.3f * px, 0.3f * py, -1.5f}, 1e-3f, {1.1f * px, 1.1f * py, 1.f}, 1e9f};
    }

    auto bytes = byte_buffer_type();
    auto vbuf = d->create_buffer(&bytes, verts.size() * 4, nullptr);
    auto ibuf = d->create_buffer(&bytes, idx.size() * 4, nullptr);
    auto rbuf = d->create_buffer(&bytes, N * sizeof(oracle_ray), nullptr);
    auto hbuf = d->create_buffer(&bytes, N * sizeof(oracle_hit), nullptr);
    auto obuf = d->create_buffer(&bytes, N * 4, nullptr);
    CHECK(vbuf.valid() && vbuf.total_size_bytes == verts.size() * 4 && vbuf.element_stride == 0);  // Type::Void: stride = ty.size() = 0, total = count bytes (cpu/mod.rs:57-81)
    AccelOption mesh_opt{};
    mesh_opt.allow_update = true;
    auto mesh = d->create_mesh(mesh_opt);
    auto accel = d->create_accel(AccelOption{});
    auto stream = d->create_stream(StreamTag::COMPUTE);
    CHECK(mesh.valid() && accel.valid() && stream.valid());

    const float xf0[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
    const float xf1[12] = {0.8f, -0.6f, 0, 0.05f, 0.6f, 0.8f, 0, -0.02f, 0, 0, 1, 0.4f};
    auto mods = [&](bool second_visible) {
        luisa::vector<AccelBuildCommand::Modification> m;
        for (uint32_t i = 0; i < 2; i++) {
            AccelBuildCommand::Modification x{i};
            x.set_primitive(mesh.handle);
            x.set_transform_data(i == 0 ? xf0 : xf1);
            x.set_visibility(i == 1 && !second_visible ? 0x00 : 0xff);
            x.set_opaque(true);
            x.set_user_id(100 + i);
            m.push_back(x);
        }
        return m;
    };

    oracle_scene *os = oracle_scene_new();
    uint64_t om = oracle_mesh_new(os);
    std::vector<oracle_hit> want(N), got(N);
    std::vector<uint32_t> want_any(N), got_any(N);
    std::atomic<int> callbacks{0};

    for (int round = 0; round < 2; round++) {
        // round 0: FORCE_BUILD of everything; round 1: moved vertices, PREFER_UPDATE, instance 1 made invisible
        if (round == 1) fill_vertices(1.3f);
        auto list = CommandList::create();
        list << luisa::make_unique<BufferUploadCommand>(vbuf.handle, 0, verts.size() * 4, verts.data());
        if (round == 0) {
            list << luisa::make_unique<BufferUploadCommand>(ibuf.handle, 0, idx.size() * 4, idx.data())
                 << luisa::make_unique<BufferUploadCommand>(rbuf.handle, 0, N * sizeof(oracle_ray), rays.data());
        }
        list << luisa::make_unique<MeshBuildCommand>(mesh.handle, round == 0 ? AccelBuildRequest::FORCE_BUILD : AccelBuildRequest::PREFER_UPDATE,
                                                      vbuf.handle, 0, verts.size() * 4, 12, ibuf.handle, 0, idx.size() * 4)
             << luisa::make_unique<AccelBuildCommand>(accel.handle, 2u, AccelBuildRequest::FORCE_BUILD, mods(round == 0), false);
        list.add_callback([&callbacks] { callbacks++; });
        d->dispatch(stream.handle, std::move(list));
        trace_closest(cdev, {stream.handle}, {accel.handle}, {rbuf.handle}, 0, {hbuf.handle}, 0, N, 0xffu);
        trace_any(cdev, {stream.handle}, {accel.handle}, {rbuf.handle}, 0, {obuf.handle}, 0, N, 0xffu);
        auto back = CommandList::create();
        back << luisa::make_unique<BufferDownloadCommand>(hbuf.handle, 0, N * sizeof(oracle_hit), got.data())
             << luisa::make_unique<BufferDownloadCommand>(obuf.handle, 0, N * 4, got_any.data());
        back.add_callback([&callbacks] { callbacks++; });
        d->dispatch(stream.handle, std::move(back));
        d->synchronize_stream(stream.handle);
        CHECK(callbacks.load() == 2 * (round + 1));

        oracle_mesh_set(os, om, verts.data(), 12, verts.size() / 3, idx.data(), 12, n_tri);
        oracle_mesh_commit(os, om);
        auto cm = mods(round == 0);
        std::vector<oracle_mod> omods;
        for (auto &&m : cm) {
            oracle_mod x{m.index, m.user_id, m.flags, m.vis_mask, om, {}};
            std::memcpy(x.affine, m.affine, sizeof(x.affine));
            omods.push_back(x);
        }
        oracle_accel_update(os, 2, omods.data(), omods.size());
        oracle_trace_closest(os, rays.data(), N, 0xffu, want.data(), 1, 0);
        oracle_trace_any(os, rays.data(), N, 0xffu, want_any.data(), 1, 0);
        uint32_t n_hit = 0, n_inst1 = 0;
        for (uint32_t i = 0; i < N; i++) {
            if (std::memcmp(&got[i], &want[i], 20) != 0 || got_any[i] != want_any[i]) {
                std::fprintf(stderr, "round %d ray %u: device {%u %u %a %a %a | %u}  oracle {%u %u %a %a %a | %u}\n", round, i, got[i].inst, got[i].prim,
                             got[i].u, got[i].v, got[i].t, got_any[i], want[i].inst, want[i].prim, want[i].u, want[i].v, want[i].t, want_any[i]);
                return 1;
            }
            n_hit += got[i].inst != ~0u;
            n_inst1 += got[i].inst == 1u;
        }
        CHECK(n_hit > N / 4);
        CHECK(round == 0 ? n_inst1 > 0 : n_inst1 == 0);
        std::printf("round %d: %u rays, %u hits (%u on instance 1), closest + any identical to the oracle\n", round, N, n_hit, n_inst1);
    }

    // a DSL kernel through the C++ interface: an ir::KernelModule made of the reference's own IR records (ir_ref_build.hpp) ->
    // DeviceInterface::create_shader(option, const ir::KernelModule *) -> the reference's ComputeDispatchCmdEncoder packs {buffer, uniform}
    // into a ShaderDispatchCommand -> the adapter turns the argument buffer into lcb_argument records (uniform bytes handed out in place)
    {
        ShaderOption so{};
        so.enable_fast_math = false;
        auto shader = d->create_shader(so, ir_ref::build_axpy_module());
        CHECK(shader.valid() && shader.block_size.x == 128u && shader.block_size.y == 1u);
        constexpr uint32_t M = 1000, LIMIT = 777;
        std::vector<float> xs(M), ys(M, -1.f);
        for (uint32_t i = 0; i < M; i++) xs[i] = 0.25f * (float)i - 3.f;
        auto fb = d->create_buffer(&bytes, M * 4, nullptr);
        uint32_t limit = LIMIT;
        ComputeDispatchCmdEncoder enc{shader.handle, 2, 16};
        enc.encode_buffer(fb.handle, 0, M * 4);
        enc.encode_uniform(&limit, sizeof(limit));
        enc.set_dispatch_size(make_uint3(M, 1u, 1u));
        auto list = CommandList::create();
        list << luisa::make_unique<BufferUploadCommand>(fb.handle, 0, M * 4, xs.data())
             << std::move(enc).build()
             << luisa::make_unique<BufferDownloadCommand>(fb.handle, 0, M * 4, ys.data());
        limit = 0;   // the command owns a copy of the uniform bytes: changing the source after encoding must not matter
        d->dispatch(stream.handle, std::move(list));
        d->synchronize_stream(stream.handle);
        for (uint32_t i = 0; i < M; i++) {
            const float want = i < LIMIT ? xs[i] * 2.0f + 1.0f : xs[i];
            if (ys[i] != want) { std::fprintf(stderr, "shader dispatch: a[%u] = %a, expected %a\n", i, ys[i], want); return 1; }
        }
        d->destroy_shader(shader.handle);
        d->destroy_buffer(fb.handle);
        std::printf("shader: reference-built ir::KernelModule -> create_shader -> ShaderDispatchCommand{buffer, uniform}: %u of %u elements updated, rest untouched\n", LIMIT, M);
    }

    // textures and bindless arrays through the C++ commands: upload -> device -> download of a Float4 image and of an RGBA8 image,
    // a BindlessArrayUpdateCommand that emplaces a buffer + a sampled texture and removes them again
    {
        constexpr uint32_t TW = 8, TH = 4;
        auto tex = d->create_texture(PixelFormat::RGBA32F, 2u, TW, TH, 1u, 1u, false, false);
        auto tex8 = d->create_texture(PixelFormat::RGBA8UNorm, 2u, TW, TH, 1u, 1u, false, false);
        CHECK(tex.valid() && tex8.valid());
        std::vector<float> src(TW * TH * 4), dst(TW * TH * 4, -1.f);
        std::vector<uint8_t> src8(TW * TH * 4), dst8(TW * TH * 4, 0xEE);
        for (size_t i = 0; i < src.size(); i++) { src[i] = 0.5f * (float)i - 7.f; src8[i] = (uint8_t)(i * 37u + 11u); }
        auto heap = d->create_bindless_array(4);
        CHECK(heap.valid());
        using Mod = BindlessArrayUpdateCommand::Modification;
        luisa::vector<Mod> emplace, remove;
        emplace.emplace_back(2u, Mod::Buffer::emplace(vbuf.handle, 0u), Mod::Texture::emplace(tex.handle, Sampler{Sampler::Filter::LINEAR_POINT, Sampler::Address::REPEAT}), Mod::Texture{});
        remove.emplace_back(2u, Mod::Buffer::remove(), Mod::Texture::remove(), Mod::Texture{});
        auto list = CommandList::create();
        list << luisa::make_unique<TextureUploadCommand>(tex.handle, PixelStorage::FLOAT4, 0u, make_uint3(TW, TH, 1u), src.data())
             << luisa::make_unique<TextureUploadCommand>(tex8.handle, PixelStorage::BYTE4, 0u, make_uint3(TW, TH, 1u), src8.data())
             << luisa::make_unique<BindlessArrayUpdateCommand>(heap.handle, std::move(emplace))
             << luisa::make_unique<TextureDownloadCommand>(tex.handle, PixelStorage::FLOAT4, 0u, make_uint3(TW, TH, 1u), dst.data())
             << luisa::make_unique<TextureDownloadCommand>(tex8.handle, PixelStorage::BYTE4, 0u, make_uint3(TW, TH, 1u), dst8.data())
             << luisa::make_unique<BindlessArrayUpdateCommand>(heap.handle, std::move(remove));
        d->dispatch(stream.handle, std::move(list));
        d->synchronize_stream(stream.handle);
        CHECK(std::memcmp(src.data(), dst.data(), src.size() * 4) == 0);
        CHECK(std::memcmp(src8.data(), dst8.data(), src8.size()) == 0);
        d->destroy_bindless_array(heap.handle);
        d->destroy_texture(tex.handle);
        d->destroy_texture(tex8.handle);
        std::printf("textures: Float4 and RGBA8 images round-trip through TextureUpload / TextureDownload; bindless emplace + remove accepted\n");
    }

    // events through the C++ interface
    auto ev = d->create_event();
    d->signal_event(ev.handle, stream.handle, 7);
    d->synchronize_event(ev.handle, 7);
    CHECK(d->is_event_completed(ev.handle, 7));
    d->destroy_event(ev.handle);

    oracle_scene_free(os);
    d->destroy_stream(stream.handle);
    d->destroy_accel(accel.handle);
    d->destroy_mesh(mesh.handle);
    for (auto h : {vbuf.handle, ibuf.handle, rbuf.handle, hbuf.handle, obuf.handle}) d->destroy_buffer(h);
    std::printf("cpp_host_check ok\n");
    return 0;
}
