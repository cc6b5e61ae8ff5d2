"""The DSL kernels of the reference's ray-tracing examples, built as ir::KernelModule values (the form `create_shader` receives).

With no Rust toolchain in the image the `track!` closures of luisa_compute/examples/{raytracing,path_tracer}.rs cannot be run
through the real frontend, so they are transcribed here statement by statement onto `ir.KernelBuilder`, emitting for every DSL
operation the `Func` / `Instruction` the frontend emits (lang/types/*.rs operators -> Func::Add.., `Var::load/store` ->
Load / Update, `if` -> If + Phi, `while` / `for` -> GenericLoop, `Callable::new_static` -> Func::Callable, struct / vector
constructors -> Func::Struct / Vec3, `lc_info!`-free).  The device lowers them with the same code path a Rust-built module
would take (csrc/ir_lower.cpp): resources arrive as captures and arguments exactly as in the examples.
"""
import math

import numpy as np

from . import ir
from .ir import Func


def common_types(k):
    """Ray / SurfaceHit / Index as the IR sees them (rtx.rs:329-366: Ray is #[repr(C, align(16))], hit structs align(8))."""
    f3 = k.array(k.f32, 3)
    ray = k.struct([f3, k.f32, f3, k.f32], align=16)
    hit = k.struct([k.u32, k.u32, k.f322, k.f32], align=8)
    return f3, ray, hit


def _to_array3(k, v, f3):
    """Expr::<[f32; 3]>::from(Float3)"""
    return k.call(Func.Array, [v.x, v.y, v.z], f3)


def _to_float3(k, a):
    """Float3::from([f32; 3])"""
    return k.vec(k.f323, a.extract(0), a.extract(1), a.extract(2))


def make_ray(k, ray_ty, f3, o, d, tmin, tmax):
    return k.make_struct(ray_ty, _to_array3(k, o, f3), k.lit(k.f32, tmin), _to_array3(k, d, f3), k.lit(k.f32, tmax))


def offset_ray_origin(k, p, n):
    """rtx.rs:517-535"""
    origin, float_scale, int_scale = 1.0 / 32.0, 1.0 / 65536.0, 256.0
    of_i = (k.f(int_scale) * n).cast(k.i323)
    p_i = p.bitcast(k.i323) + p.lt(0.0).select(-of_i, of_i)
    return p.abs().lt(origin).select(p + k.f(float_scale) * n, p_i.bitcast(k.f323))


def raytracing_kernel(accel_handle, image_handle, width, height):
    """examples/raytracing.rs:44-71 — `Kernel::<fn()>`: the accel and the Byte4 image are captures."""
    k = ir.KernelBuilder(block_size=(16, 16, 1))
    f3, ray_ty, hit_ty = common_types(k)
    accel = k.capture_accel(accel_handle)
    img = k.capture_tex2d(k.f324, image_handle)

    def body():
        px = k.dispatch_id().permute(0, 1)
        xy = px.cast(k.f322) / k.vec(k.f322, float(width), float(height))
        xy = k.f(2.0) * xy - 1.0
        o = k.vec(k.f323, 0.0, 0.0, -1.0)
        d = (k.vec(k.f323, xy.x, xy.y, 0.0) - o).normalize()
        ray = make_ray(k, ray_ty, f3, o, d, 1e-3, 1e9)
        hit = accel.trace_closest(ray, 0xFF, hit_ty)
        bary = hit.extract(2)
        color = hit.extract(0).ne(0xFFFFFFFF).select(k.vec(k.f323, bary.x, bary.y, 1.0), k.vec(k.f323, 0.0, 0.0, 0.0))
        img.tex_write(px, k.vec(k.f324, color.x, color.y, color.z, 1.0))
    k.body(body)
    k.finish()
    return k


def raytracing_rays(width, height):
    """The rays `raytracing_kernel` generates, evaluated on the host in the kernel's fp32 operation order
    (normalize = v * (1 / sqrt(dot(v, v))), device_math.h:3588): rows of (o, tmin, d, tmax), pixel-major."""
    f = np.float32
    ys, xs = np.meshgrid(np.arange(height, dtype=np.float32), np.arange(width, dtype=np.float32), indexing="ij")
    x = f(2.0) * (xs / f(width)) - f(1.0)
    y = f(2.0) * (ys / f(height)) - f(1.0)
    dx, dy, dz = x - f(0.0), y - f(0.0), np.full_like(x, f(0.0) - f(-1.0))
    inv = f(1.0) / np.sqrt((dx * dx + dy * dy) + dz * dz)
    rays = np.zeros((height * width, 8), np.float32)
    rays[:, 2] = -1.0
    rays[:, 3] = f(1e-3)
    rays[:, 4:7] = np.stack([dx * inv, dy * inv, dz * inv], -1).reshape(-1, 3)
    rays[:, 7] = f(1e9)
    return rays


def trace_buffer_kernel(any_hit=False, mask=0xFF, block_size=(64, 1, 1)):
    """The smallest DSL kernel on the hot path — what a luisa-compute-rs user writes to trace a buffer of rays (config C3):

        Kernel::<fn(Buffer<Ray>, Buffer<SurfaceHit>, Accel)>::new(&device, &track!(|rays, hits, accel| {
            let i = dispatch_id().x;
            hits.write(i, accel.intersect(rays.read(i), mask));      // rtx.rs:774-795 -> Func::RayTracingTraceClosest
        }))

    (`any_hit`: Buffer<bool> stored as u32, accel.intersect_any -> Func::RayTracingTraceAny, rtx.rs:796-817).  Block size: the frontend's
    default for kernels that do not call set_block_size (runtime/kernel.rs:506)."""
    k = ir.KernelBuilder(block_size=block_size)
    _, ray_ty, hit_ty = common_types(k)
    rays = k.arg_buffer(ray_ty)
    out = k.arg_buffer(k.u32 if any_hit else hit_ty)
    accel = k.arg_accel()

    def body():
        i = k.dispatch_id().x
        ray = rays.read(i)
        if any_hit:
            out.write(i, accel.trace_any(ray, mask).select(k.u(1), k.u(0)))
        else:
            out.write(i, accel.trace_closest(ray, mask, hit_ty))
    k.body(body)
    k.finish()
    return k


CBOX_MATERIALS = [(0.725, 0.710, 0.680)] * 3 + [(0.140, 0.450, 0.091), (0.630, 0.065, 0.050)] + [(0.725, 0.710, 0.680)] * 2 + [(0.0, 0.0, 0.0)]
SPP_PER_DISPATCH = 32
FRAC_1_PI = float(np.float32(0.318309886183790671537767526745028724))
F32_MAX = float(np.finfo(np.float32).max)
TAN_HALF_FOV = float(np.float32(np.tan(np.float32(0.5) * (np.float32(27.8) * np.float32(np.pi) / np.float32(180.0)))))


CUTOUT_STRIPES = {5: (6.0, 0.6), 6: (5.0, 0.5)}   # path_tracer_cutout.rs:364-373: instance -> (frequency, kept fraction) of bary.y stripes


def path_tracer_kernel(vertex_heap_handle, index_heap_handle, spp_per_dispatch=SPP_PER_DISPATCH, max_depth=10, polynomial_sincos=False, cutout=False):
    """examples/path_tracer.rs:247-455 — `Kernel::<fn(Tex2d<Float4>, Tex2d<u32>, Accel, Uint2)>`; the two bindless heaps are captures.

    cutout=True is examples/path_tracer_cutout.rs: `accel.traverse(ray).on_surface_hit(|c| if filter(&c) { c.commit() }).trace()` replaces
    intersect, `traverse_any` replaces intersect_any (:379-389, :435-445); the filter cuts bary.y stripes out of the two boxes (:364-373).
    The instances must then be pushed non-opaque (:250) — opaque instances never reach the callback.

    polynomial_sincos=False emits Func::Sin / Func::Cos for the hemisphere sample like the example.  True replaces them by a
    callable evaluating the fixed polynomial of csrc/path_tracer.cu / oracle.c (sincos_2pi), so that the random walks are
    bit-reproducible against the CPU restatement (libm and libdevice differ in the last ulp of sin / cos)."""
    k = ir.KernelBuilder(block_size=(16, 16, 1))
    f3, ray_ty, hit_ty = common_types(k)
    index_ty = k.array(k.u32, 3)
    onb_ty = k.struct([k.f323, k.f323, k.f323])
    vertex_heap = k.capture_bindless(vertex_heap_handle)
    index_heap = k.capture_bindless(index_heap_handle)
    image = k.arg_tex2d(k.f324)
    seed_image = k.arg_tex2d(k.u32)
    accel = k.arg_accel()
    resolution = k.arg_uniform(k.u322)

    def lcg_body(state):  # path_tracer.rs:271-279
        state.store(k.u(1664525) * state.load() + k.u(1013904223))
        k.return_((state.load() & k.u(0x00FFFFFF)).cast(k.f32) * k.f(1.0 / 16777216.0))
    lcg = k.callable([(k.u32, False)], k.f32, lcg_body)

    def sincos_body(u, s_out, c_out):  # csrc/path_tracer.cu sincos_2pi, in IR operations
        kf = (u * 4.0 + 0.5).floor()
        r = u - kf * 0.25
        x = r * k.f(6.28318530717958647692)
        x2 = x * x
        sp = x2.fma(k.f(2.7557319e-6), k.f(-1.9841270e-4))
        sp = sp.fma(x2, k.f(8.3333333e-3))
        sp = sp.fma(x2, k.f(-1.6666667e-1))
        sp = (sp * x2).fma(x, x)
        cp = x2.fma(k.f(2.4801587e-5), k.f(-1.3888889e-3))
        cp = cp.fma(x2, k.f(4.1666667e-2))
        cp = cp.fma(x2, k.f(-0.5))
        cp = cp.fma(x2, k.f(1.0))
        q = kf.cast(k.i32) & k.i(3)
        s_out.store(q.eq(0).select(sp, q.eq(1).select(cp, q.eq(2).select(-sp, -cp))))
        c_out.store(q.eq(0).select(cp, q.eq(1).select(-sp, q.eq(2).select(-cp, sp))))
        k.return_()
    sincos = k.callable([(k.f32, True), (k.f32, False), (k.f32, False)], k.void, sincos_body) if polynomial_sincos else None

    committed_ty = k.struct([k.u32, k.u32, k.f322, k.u32, k.f32], align=8)

    def cutout_query(ray_value, any_hit):
        """RayQueryAll / RayQueryAny with the example's filter as the on_surface_hit block; returns the CommittedHit."""
        rq = accel.query(ray_value, k.u(0xFF), any_hit)

        def on_surface_hit():
            cand = k.call(Func.RayQueryTriangleCandidateHit, [rq], hit_ty)
            c_inst, v = cand.extract(0), cand.extract(2).y
            valid = None
            for inst_id, (freq, keep) in sorted(CUTOUT_STRIPES.items()):
                ok = c_inst.ne(inst_id) | (v * k.f(freq)).unary(Func.Fract).lt(keep)
                valid = ok if valid is None else (valid & ok)
            k.if_(valid, lambda: k.call(Func.RayQueryCommitTriangle, [rq], k.void))
        k.ray_query(rq, on_surface_hit, None)
        return k.call(Func.RayQueryCommittedHit, [rq], committed_ty)

    def body():
        materials = k.const(k.array(k.f323, 8), CBOX_MATERIALS)
        coord = k.dispatch_id().permute(0, 1)
        frame_size = k.min(resolution.x, resolution.y).cast(k.f32)
        state = k.local_zero(k.u32)
        state.store(seed_image.tex_read(coord))
        rx = lcg(state)
        ry = lcg(state)
        pixel = (coord.cast(k.f322) + k.vec(k.f322, rx, ry)) / frame_size * 2.0 - 1.0
        radiance = k.local_zero(k.f323)
        radiance.store(k.vec(k.f323, 0.0, 0.0, 0.0))
        sample = k.local_zero(k.u32)

        def sample_body():
            p = pixel * k.vec(k.f322, 1.0, -1.0)
            origin = k.vec(k.f323, -0.01, 0.995, 5.0)
            pixel3 = origin + k.vec(k.f323, p.x * k.f(TAN_HALF_FOV), p.y * k.f(TAN_HALF_FOV), -1.0)
            direction = (pixel3 - origin).normalize()
            ray = k.local_zero(ray_ty)
            ray.store(make_ray(k, ray_ty, f3, origin, direction, 0.0, F32_MAX))
            beta = k.local_zero(k.f323)
            beta.store(k.vec(k.f323, 1.0, 1.0, 1.0))
            pdf_bsdf = k.local_zero(k.f32)
            pdf_bsdf.store(k.f(0.0))
            light_position = k.vec(k.f323, -0.24, 1.98, 0.16)
            light_u = k.vec(k.f323, -0.24, 1.98, -0.22) - light_position
            light_v = k.vec(k.f323, 0.23, 1.98, 0.16) - light_position
            light_emission = k.vec(k.f323, 17.0, 12.0, 4.0)
            light_area = light_u.cross(light_v).length()
            light_normal = light_u.cross(light_v).normalize()
            depth = k.local_zero(k.u32)

            def bounce():
                hit = cutout_query(ray.load(), False) if cutout else accel.trace_closest(ray.load(), 0xFF, hit_ty)
                inst, prim, bary = hit.extract(0), hit.extract(1), hit.extract(2)   # CommittedHit and SurfaceHit agree on these fields
                k.if_(inst.ne(0xFFFFFFFF).not_(), lambda: k.break_())
                tri = index_heap.bindless_buffer_read(inst, prim, index_ty)
                p0 = _to_float3(k, vertex_heap.bindless_buffer_read(inst, tri.extract(0), f3))
                p1 = _to_float3(k, vertex_heap.bindless_buffer_read(inst, tri.extract(1), f3))
                p2 = _to_float3(k, vertex_heap.bindless_buffer_read(inst, tri.extract(2), f3))
                pnt = (k.f(1.0) - bary.x - bary.y) * p0 + bary.x * p1 + bary.y * p2  # SurfaceHit::interpolate, rtx.rs:384
                n = (p1 - p0).cross(p2 - p0).normalize()
                origin_w = _to_float3(k, ray.gep(0).load())
                direction_w = _to_float3(k, ray.gep(2).load())
                cos_wi = -direction_w.dot(n)
                k.if_(cos_wi.lt(1e-4), lambda: k.break_())
                pp = offset_ray_origin(k, pnt, n)
                albedo = materials.extract(inst)

                def hit_light():
                    def first():
                        radiance.store(radiance.load() + light_emission)

                    def later():
                        pdf_light = (pnt - origin_w).length_squared() / (light_area * cos_wi)
                        mis_weight = pdf_bsdf.load() / k.max(pdf_bsdf.load() + pdf_light, k.f(1e-4))
                        radiance.store(radiance.load() + mis_weight * beta.load() * light_emission)
                    k.if_(depth.load().eq(0), first, later)
                    k.break_()

                def sample_light():
                    ux_light = lcg(state)
                    uy_light = lcg(state)
                    p_light = light_position + ux_light * light_u + uy_light * light_v
                    pp_light = offset_ray_origin(k, p_light, light_normal)
                    d_light = (pp - pp_light).length()
                    wi_light = (pp_light - pp).normalize()
                    shadow_ray = make_ray(k, ray_ty, f3, offset_ray_origin(k, pp, n), wi_light, 0.0, d_light)
                    occluded = cutout_query(shadow_ray, True).extract(3).ne(0) if cutout else accel.trace_any(shadow_ray, 0xFF)
                    cos_wi_light = wi_light.dot(n)
                    cos_light = -light_normal.dot(wi_light)

                    def add_direct():
                        pdf_light = (d_light * d_light) / (light_area * cos_light)
                        pdf_b = cos_wi_light * k.f(FRAC_1_PI)
                        mis_weight = pdf_light / k.max(pdf_light + pdf_b, k.f(1e-4))
                        bsdf = albedo * k.f(FRAC_1_PI) * cos_wi_light
                        radiance.store(radiance.load() + beta.load() * bsdf * mis_weight * light_emission / k.max(pdf_light, k.f(1e-4)))
                    k.if_(occluded.not_() & cos_wi_light.gt(1e-4) & cos_light.gt(1e-4), add_direct)
                k.if_(inst.eq(7), hit_light, sample_light)
                # sample BSDF: make_onb (path_tracer.rs:311-321) + cosine_sample_hemisphere (:323-327)
                binormal = k.if_phi(n.x.abs().gt(n.z.abs()), lambda: k.vec(k.f323, -n.y, n.x, 0.0), lambda: k.vec(k.f323, 0.0, -n.z, n.y))
                tangent = binormal.cross(n).normalize()
                onb = k.make_struct(onb_ty, tangent, binormal, n)
                ux = lcg(state)
                uy = lcg(state)
                r = ux.sqrt()
                if polynomial_sincos:
                    s_var, c_var = k.local_zero(k.f32), k.local_zero(k.f32)
                    sincos(uy, s_var, c_var)
                    sin_phi, cos_phi = s_var.load(), c_var.load()
                else:
                    phi = k.f(2.0 * math.pi) * uy
                    sin_phi, cos_phi = phi.sin(), phi.cos()
                local = k.vec(k.f323, r * cos_phi, r * sin_phi, (k.f(1.0) - ux).sqrt())
                new_direction = onb.extract(0) * local.x + onb.extract(1) * local.y + onb.extract(2) * local.z  # Onb::to_world
                ray.store(make_ray(k, ray_ty, f3, pp, new_direction, 0.0, F32_MAX))
                beta.store(beta.load() * albedo)
                pdf_bsdf.store(cos_wi * k.f(FRAC_1_PI))
                # russian roulette
                lum = k.vec(k.f323, 0.212671, 0.715160, 0.072169).dot(beta.load())
                k.if_(lum.eq(0.0), lambda: k.break_())
                q = k.max(lum, k.f(0.05))
                rr = lcg(state)
                k.if_(rr.gt(q), lambda: k.break_())
                beta.store(beta.load() / q)
                depth.store(depth.load() + k.u(1))
            k.generic_loop(lambda: depth.load().lt(max_depth), bounce)
        k.generic_loop(lambda: sample.load().lt(spp_per_dispatch), sample_body, lambda: sample.store(sample.load() + k.u(1)))
        radiance.store(radiance.load() / k.f(float(spp_per_dispatch)))
        seed_image.tex_write(coord, state.load())
        k.if_(radiance.load().is_nan().any(), lambda: radiance.store(k.vec(k.f323, 0.0, 0.0, 0.0)))
        clamped = radiance.load().clamp(k.vec(k.f323, k.f(0.0)), k.vec(k.f323, k.f(30.0)))
        old = image.tex_read(coord)
        spp = old.w
        total = clamped + old.permute(0, 1, 2)
        image.tex_write(coord, k.vec(k.f324, total.x, total.y, total.z, spp + 1.0))
    k.body(body)
    k.finish()
    return k


def display_kernel():
    """examples/path_tracer.rs:465-479 — `Kernel::<fn(Tex2d<Float4>, Tex2d<Float4>)>`: accumulated radiance / spp through the sRGB transfer
    curve into the display image (whose storage is the swapchain's, Byte4: the texel conversion rounds to 8 bits, cpu_texture.h)."""
    k = ir.KernelBuilder(block_size=(16, 16, 1))
    acc = k.arg_tex2d(k.f324)
    display = k.arg_tex2d(k.f324)

    def body():
        coord = k.dispatch_id().permute(0, 1)
        texel = acc.tex_read(coord)
        radiance = texel.permute(0, 1, 2) / texel.w
        e = 1.0 / 2.4
        r = k.f(1.055) * k.call(Func.Powf, [radiance, k.vec(k.f323, e, e, e)], k.f323) - 0.055   # the scalar exponent is spread to the vector (ops/spread.rs:367)
        srgb = radiance.lt(0.0031308).select(radiance * 12.92, r)
        display.tex_write(coord, k.vec(k.f324, srgb.x, srgb.y, srgb.z, 1.0))
    k.body(body)
    k.finish()
    return k


def ray_query_kernel(radius, any_hit=False):
    """The query of examples/ray_query.rs:135-170 over a ray buffer: `accel.traverse(ray).on_surface_hit(|c| if disc(c.bary) { c.commit() })`,
    one ray per thread, CommittedHit written per ray.  Args: rays Buffer<Ray>, committed Buffer<CommittedHit>, accel, mask: u32."""
    k = ir.KernelBuilder(block_size=(128, 1, 1))
    f3, ray_ty, hit_ty = common_types(k)
    committed_ty = k.struct([k.u32, k.u32, k.f322, k.u32, k.f32], align=8)
    rays, out, accel, mask = k.arg_buffer(ray_ty), k.arg_buffer(committed_ty), k.arg_accel(), k.arg_uniform(k.u32)

    def body():
        i = k.dispatch_id().x
        rq = accel.query(rays.read(i), mask, any_hit)

        def on_surface_hit():
            cand = k.call(Func.RayQueryTriangleCandidateHit, [rq], hit_ty)
            bary = cand.extract(2)
            u, v = bary.x, bary.y
            w = k.f(1.0) - u - v
            r2 = k.f(radius) * k.f(radius)
            xy = w.fma(w, u * u)
            yz = u.fma(u, v * v)
            xz = w.fma(w, v * v)
            k.if_(xy.lt(r2) & yz.lt(r2) & xz.lt(r2), lambda: k.call(Func.RayQueryCommitTriangle, [rq], k.void))
        k.ray_query(rq, on_surface_hit, None)
        out.write(i, k.call(Func.RayQueryCommittedHit, [rq], committed_ty))
    k.body(body)
    k.finish()
    return k


def tiled_path_tracer_kernel(vertex_heap_handle, index_heap_handle, camera, light, n_instances, spp_per_dispatch=32, max_depth=5, tile=64, block=16, regenerate=False, tile_counters=False):
    """BASELINE config C5: the path tracer of examples/path_tracer.rs generalised to an instanced scene and to tile sharding
    (SURVEY.md §8d / §8e).  One thread per pixel of a 64x64 tile; `tiles[k]` names the global tile a rank's k-th local tile is
    (Morton round-robin, sharding.tiles_of_rank), results accumulate in a packed per-rank tile buffer that ONE all-gather
    assembles.  The random stream is keyed by the GLOBAL pixel index and the frame number, so the image does not depend on
    the number of GPUs.  Differences from the Cornell kernel: pinhole camera given by `camera` (origin, forward, right, up,
    tan_half_fov), object-space vertices go through RayTracingInstanceTransform, normals face the viewer, a constant sky, one
    emissive quad `light` = (position, u, v, emission, instance index), albedo by instance.
    Args: tiles Buffer<u32>, out Buffer<Float4>, accel, params {resolution: Uint2, frame: u32, n_local_tiles: u32},
    counters Buffer<u64> ([0] closest-hit rays, [1] any-hit rays traced, for Mrays/s; with `tile_counters` [2 + global tile id] rays traced
    for that tile).  `block`: edge of the square thread block.
    `regenerate`: one loop whose iteration is ONE bounce — a lane whose path ended starts its next sample at once instead of idling
    until the longest path of the warp's current sample is done (the nested sample / bounce loops of the example leave 3 of 4 lanes
    idle on C2, profiles/r01u_*).  Each pixel still consumes its random stream in the same order, so the image is bit-identical either
    way.  Off by default: on C5 paths average 1.13 rays per sample, so there is little idling to recover and the loss of ray coherence
    inside a warp costs more (300.9 vs 284.1 ms, profiles/r01v_c5_regeneration.txt)."""
    k = ir.KernelBuilder(block_size=(block, block, 1))
    f3, ray_ty, hit_ty = common_types(k)
    index_ty = k.array(k.u32, 3)
    params_ty = k.struct([k.u322, k.u32, k.u32])
    vertex_heap = k.capture_bindless(vertex_heap_handle)
    index_heap = k.capture_bindless(index_heap_handle)
    tiles, out, accel, params = k.arg_buffer(k.u32), k.arg_buffer(k.f324), k.arg_accel(), k.arg_uniform(params_ty)
    counters = k.arg_buffer(k.u64)
    cam_o, cam_f, cam_r, cam_u, tan_half = camera
    l_pos, l_u, l_v, l_emission, l_inst = light
    palette = [(0.55, 0.50, 0.42), (0.42, 0.55, 0.40), (0.60, 0.45, 0.38), (0.45, 0.48, 0.58), (0.58, 0.56, 0.40), (0.40, 0.52, 0.52), (0.52, 0.42, 0.50), (0.50, 0.50, 0.50)]
    albedos = [palette[i % 8] for i in range(n_instances)]
    albedos[l_inst] = (0.0, 0.0, 0.0)

    def lcg_body(state):
        state.store(k.u(1664525) * state.load() + k.u(1013904223))
        k.return_((state.load() & k.u(0x00FFFFFF)).cast(k.f32) * k.f(1.0 / 16777216.0))
    lcg = k.callable([(k.u32, False)], k.f32, lcg_body)

    def hash_body(v):  # PCG output permutation (Jarzynski & Olano 2020), integer-exact on every device
        s = v * k.u(747796405) + k.u(2891336453)
        w = ((s >> ((s >> k.u(28)) + k.u(4))) ^ s) * k.u(277803737)
        k.return_((w >> k.u(22)) ^ w)
    pcg = k.callable([(k.u32, True)], k.u32, hash_body)

    def sincos_body(u, s_out, c_out):
        kf = (u * 4.0 + 0.5).floor()
        x = (u - kf * 0.25) * k.f(6.28318530717958647692)
        x2 = x * x
        sp = x2.fma(k.f(2.7557319e-6), k.f(-1.9841270e-4)).fma(x2, k.f(8.3333333e-3)).fma(x2, k.f(-1.6666667e-1))
        sp = (sp * x2).fma(x, x)
        cp = x2.fma(k.f(2.4801587e-5), k.f(-1.3888889e-3)).fma(x2, k.f(4.1666667e-2)).fma(x2, k.f(-0.5)).fma(x2, k.f(1.0))
        q = kf.cast(k.i32) & k.i(3)
        s_out.store(q.eq(0).select(sp, q.eq(1).select(cp, q.eq(2).select(-sp, -cp))))
        c_out.store(q.eq(0).select(cp, q.eq(1).select(-sp, q.eq(2).select(-cp, sp))))
        k.return_()
    sincos = k.callable([(k.f32, True), (k.f32, False), (k.f32, False)], k.void, sincos_body)

    def body():
        materials = k.const(k.array(k.f323, n_instances), albedos)
        res = params.extract(0)
        frame = params.extract(1)
        did = k.dispatch_id()
        lx, gy = did.x, did.y
        local_tile, ly = gy / k.u(tile), gy % k.u(tile)
        tiles_x = (res.x + k.u(tile - 1)) / k.u(tile)
        tile_id = tiles.read(local_tile)
        px = (tile_id % tiles_x) * k.u(tile) + lx
        py = (tile_id / tiles_x) * k.u(tile) + ly
        k.if_((px.ge(res.x) | py.ge(res.y)), lambda: k.return_())
        slot = local_tile * k.u(tile * tile) + ly * k.u(tile) + lx
        state = k.local_zero(k.u32)
        state.store(pcg(pcg(py * res.x + px) + frame))
        radiance = k.local_zero(k.f323)
        sample = k.local_zero(k.u32)
        n_closest, n_any = k.local_zero(k.u32), k.local_zero(k.u32)
        sky = k.vec(k.f323, 0.20, 0.25, 0.32)
        light_position = k.vec(k.f323, *l_pos)
        light_u, light_v = k.vec(k.f323, *l_u), k.vec(k.f323, *l_v)
        light_emission = k.vec(k.f323, *l_emission)
        light_area = light_u.cross(light_v).length()
        light_normal = light_u.cross(light_v).normalize()
        aspect = res.x.cast(k.f32) / res.y.cast(k.f32)

        ray = k.local_zero(ray_ty)
        beta = k.local_zero(k.f323)
        pdf_bsdf = k.local_zero(k.f32)
        depth = k.local_zero(k.u32)
        alive = k.local(k.b(False))

        def end_path():
            if regenerate:
                alive.store(k.b(False))
                k.continue_()
            else:
                k.break_()

        def start_path():
            jx, jy = lcg(state), lcg(state)
            sx = ((px.cast(k.f32) + jx) / res.x.cast(k.f32) * 2.0 - 1.0) * k.f(tan_half) * aspect
            sy = (k.f(1.0) - (py.cast(k.f32) + jy) / res.y.cast(k.f32) * 2.0) * k.f(tan_half)
            origin = k.vec(k.f323, *cam_o)
            direction = (k.vec(k.f323, *cam_f) + sx * k.vec(k.f323, *cam_r) + sy * k.vec(k.f323, *cam_u)).normalize()
            ray.store(make_ray(k, ray_ty, f3, origin, direction, 1e-4, F32_MAX))
            beta.store(k.vec(k.f323, 1.0, 1.0, 1.0))
            pdf_bsdf.store(k.f(0.0))
            depth.store(k.u(0))

        if True:
            def bounce():
                hit = accel.trace_closest(ray.load(), 0xFF, hit_ty)
                n_closest.store(n_closest.load() + k.u(1))
                inst, prim, bary = hit.extract(0), hit.extract(1), hit.extract(2)

                def missed():
                    radiance.store(radiance.load() + beta.load() * sky)
                    end_path()
                k.if_(inst.eq(0xFFFFFFFF), missed)
                tri = index_heap.bindless_buffer_read(inst, prim, index_ty)
                xf = k.call(Func.RayTracingInstanceTransform, [accel, inst], k.matrix(4))

                def world(i):
                    p = _to_float3(k, vertex_heap.bindless_buffer_read(inst, tri.extract(i), f3))
                    return (xf * k.vec(k.f324, p.x, p.y, p.z, 1.0)).permute(0, 1, 2)
                p0, p1, p2 = world(0), world(1), world(2)
                pnt = (k.f(1.0) - bary.x - bary.y) * p0 + bary.x * p1 + bary.y * p2
                ng = (p1 - p0).cross(p2 - p0).normalize()
                origin_w = _to_float3(k, ray.gep(0).load())
                direction_w = _to_float3(k, ray.gep(2).load())
                n = direction_w.dot(ng).gt(0.0).select(-ng, ng)
                cos_wi = -direction_w.dot(n)
                k.if_(cos_wi.lt(1e-4), end_path)
                pp = offset_ray_origin(k, pnt, n)
                albedo = materials.extract(inst)

                def hit_light():
                    def first():
                        radiance.store(radiance.load() + light_emission)

                    def later():
                        pdf_light = (pnt - origin_w).length_squared() / (light_area * cos_wi)
                        mis_weight = pdf_bsdf.load() / k.max(pdf_bsdf.load() + pdf_light, k.f(1e-4))
                        radiance.store(radiance.load() + mis_weight * beta.load() * light_emission)
                    k.if_(depth.load().eq(0), first, later)
                    end_path()

                def sample_light():
                    p_light = light_position + lcg(state) * light_u + lcg(state) * light_v
                    pp_light = offset_ray_origin(k, p_light, light_normal)
                    d_light = (pp - pp_light).length()
                    wi_light = (pp_light - pp).normalize()
                    shadow_ray = make_ray(k, ray_ty, f3, offset_ray_origin(k, pp, n), wi_light, 0.0, d_light)
                    occluded = accel.trace_any(shadow_ray, 0xFF)
                    n_any.store(n_any.load() + k.u(1))
                    cos_wi_light = wi_light.dot(n)
                    cos_light = -light_normal.dot(wi_light)

                    def add_direct():
                        pdf_light = (d_light * d_light) / (light_area * cos_light)
                        pdf_b = cos_wi_light * k.f(FRAC_1_PI)
                        mis_weight = pdf_light / k.max(pdf_light + pdf_b, k.f(1e-4))
                        bsdf = albedo * k.f(FRAC_1_PI) * cos_wi_light
                        radiance.store(radiance.load() + beta.load() * bsdf * mis_weight * light_emission / k.max(pdf_light, k.f(1e-4)))
                    k.if_(occluded.not_() & cos_wi_light.gt(1e-4) & cos_light.gt(1e-4), add_direct)
                k.if_(inst.eq(l_inst), hit_light, sample_light)
                binormal = k.if_phi(n.x.abs().gt(n.z.abs()), lambda: k.vec(k.f323, -n.y, n.x, 0.0), lambda: k.vec(k.f323, 0.0, -n.z, n.y)).normalize()
                tangent = binormal.cross(n).normalize()
                ux, uy = lcg(state), lcg(state)
                r = ux.sqrt()
                s_var, c_var = k.local_zero(k.f32), k.local_zero(k.f32)
                sincos(uy, s_var, c_var)
                new_direction = tangent * (r * c_var.load()) + binormal * (r * s_var.load()) + n * (k.f(1.0) - ux).sqrt()
                ray.store(make_ray(k, ray_ty, f3, pp, new_direction, 0.0, F32_MAX))
                beta.store(beta.load() * albedo)
                pdf_bsdf.store(cos_wi * k.f(FRAC_1_PI))
                lum = k.vec(k.f323, 0.212671, 0.715160, 0.072169).dot(beta.load())
                k.if_(lum.eq(0.0), end_path)
                q = k.max(lum, k.f(0.05))
                k.if_(lcg(state).gt(q), end_path)
                beta.store(beta.load() / q)
                depth.store(depth.load() + k.u(1))
                if regenerate:
                    k.if_(depth.load().ge(max_depth), lambda: alive.store(k.b(False)))

        if regenerate:
            def iteration():
                def begin():
                    start_path()
                    sample.store(sample.load() + k.u(1))
                    alive.store(k.b(True))
                k.if_(alive.load().not_(), begin)
                bounce()
            k.generic_loop(lambda: sample.load().lt(spp_per_dispatch) | alive.load(), iteration)
        else:
            def sample_body():
                start_path()
                k.generic_loop(lambda: depth.load().lt(max_depth), bounce)
            k.generic_loop(lambda: sample.load().lt(spp_per_dispatch), sample_body, lambda: sample.store(sample.load() + k.u(1)))
        rad = radiance.load() / k.f(float(spp_per_dispatch))
        rad = rad.is_nan().any().select(k.vec(k.f323, 0.0, 0.0, 0.0), rad).clamp(k.vec(k.f323, k.f(0.0)), k.vec(k.f323, k.f(30.0)))
        old = out.read(slot)
        out.write(slot, k.vec(k.f324, rad.x + old.x, rad.y + old.y, rad.z + old.z, old.w + 1.0))
        counters.atomic_fetch_add(0, n_closest.load().cast(k.u64))
        counters.atomic_fetch_add(1, n_any.load().cast(k.u64))
        if tile_counters:   # rays traced per global tile: the cost map of the multi-GPU partition (sharding.balanced_bounds)
            counters.atomic_fetch_add(k.u(2) + tile_id, (n_closest.load() + n_any.load()).cast(k.u64))
    k.body(body)
    k.finish()
    return k


def sphere_query_kernel(translate):
    """examples/ray_query.rs:125-200 with an analytic sphere test in place of the example's sphere tracing loop: a RayQuery whose
    on_surface_hit commits every triangle candidate and whose on_procedural_hit intersects the candidate's sphere
    (`spheres[prim]` = centre.xyz, radius; `translate` = the procedural instance's translation, as in the example) and commits with
    its own t.  Args: rays Buffer<Ray>, committed Buffer<CommittedHit>, accel, spheres Buffer<Float4>, mask: u32."""
    k = ir.KernelBuilder(block_size=(128, 1, 1))
    f3, ray_ty, hit_ty = common_types(k)
    committed_ty = k.struct([k.u32, k.u32, k.f322, k.u32, k.f32], align=8)
    procedural_ty = k.struct([k.u32, k.u32])
    rays, out, accel, spheres, mask = k.arg_buffer(ray_ty), k.arg_buffer(committed_ty), k.arg_accel(), k.arg_buffer(k.f324), k.arg_uniform(k.u32)

    def body():
        i = k.dispatch_id().x
        rq = accel.query(rays.read(i), mask)

        def on_surface_hit():
            k.call(Func.RayQueryCommitTriangle, [rq], k.void)

        def on_procedural_hit():
            cand = k.call(Func.RayQueryProceduralCandidateHit, [rq], procedural_ty)
            ray = k.call(Func.RayQueryWorldSpaceRay, [rq], ray_ty)
            o, d = _to_float3(k, ray.extract(0)), _to_float3(k, ray.extract(2))
            s = spheres.read(cand.extract(1))
            centre = s.permute(0, 1, 2) + k.vec(k.f323, *translate)
            oc = o - centre
            a = d.dot(d)
            b = oc.dot(d)
            c = oc.dot(oc) - s.w * s.w
            disc = b * b - a * c

            def inside():
                t = (-b - disc.sqrt()) / a
                k.if_(t.ge(ray.extract(1)) & t.lt(ray.extract(3)), lambda: k.call(Func.RayQueryCommitProcedural, [rq, t], k.void))
            k.if_(disc.ge(0.0), inside)
        k.ray_query(rq, on_surface_hit, on_procedural_hit)
        out.write(i, k.call(Func.RayQueryCommittedHit, [rq], committed_ty))
    k.body(body)
    k.finish()
    return k
