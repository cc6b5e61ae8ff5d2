// path_tracer.cu — the kernel of luisa_compute/examples/path_tracer.rs:247-455, written by hand in the form the
// IR -> CUDA lowering would emit: one thread per pixel (block 16x16, set_block_size :252), the whole path loop inside the
// thread, `accel.intersect` / `accel.intersect_any` as calls to the single-ray traversal routines of trace_device.cuh —
// the role lc_trace_closest / lc_trace_any play for the CPU backend's generated code (cpu_resource.h:288-294).
// It exists so that BASELINE config C2 (Cornell box, 1024x1024, 256 spp) runs end to end on the new traversal code.
//
// Arithmetic: compiled with -fmad=false and written with one operation per expression in the example's order, so the
// CPU restatement in oracle/oracle_pt.c reproduces every sample bit for bit (sin / cos of the hemisphere sample use the
// fixed polynomial below instead of libm / libdevice, whose last-ulp differences would decorrelate the two random walks).
#include "../../include/lc_b200_api.h"
#include "trace_device.cuh"

namespace lcb {

namespace {

struct V3 { float x, y, z; };
__device__ __forceinline__ V3 v3(float x, float y, float z) { return V3{x, y, z}; }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return V3{a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return V3{a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ V3 operator*(V3 a, V3 b) { return V3{a.x * b.x, a.y * b.y, a.z * b.z}; }
__device__ __forceinline__ V3 operator*(V3 a, float s) { return V3{a.x * s, a.y * s, a.z * s}; }
__device__ __forceinline__ V3 operator*(float s, V3 a) { return V3{s * a.x, s * a.y, s * a.z}; }
__device__ __forceinline__ V3 operator/(V3 a, float s) { return V3{a.x / s, a.y / s, a.z / s}; }
__device__ __forceinline__ float dot(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
__device__ __forceinline__ V3 cross(V3 a, V3 b) { return V3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
__device__ __forceinline__ float length(V3 a) { return sqrtf(dot(a, a)); }
// lc_normalize(v) = v * rsqrt(dot(v, v)) with rsqrt(x) = 1 / sqrt(x): cpu/codegen/device_math.h:3588, cpu_prelude.h:7
__device__ __forceinline__ V3 normalize(V3 a) { return a * (1.0f / sqrtf(dot(a, a))); }

// lc/src/rtx.rs:517-535
__device__ __forceinline__ V3 offset_ray_origin(V3 p, V3 n) {
    const float origin = 1.0f / 32.0f, float_scale = 1.0f / 65536.0f, int_scale = 256.0f;
    const float pv[3] = {p.x, p.y, p.z}, nv[3] = {n.x, n.y, n.z};
    float o[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const int of_i = (int)(int_scale * nv[k]);
        const int p_i = __float_as_int(pv[k]) + (pv[k] < 0.0f ? -of_i : of_i);
        o[k] = fabsf(pv[k]) < origin ? pv[k] + float_scale * nv[k] : __int_as_float(p_i);
    }
    return V3{o[0], o[1], o[2]};
}

// sin(2 pi u), cos(2 pi u) for u in [0, 1): quadrant reduction on u itself (exact), then fixed Taylor polynomials on
// [-pi/4, pi/4] with explicit fma order.  Restated in oracle_pt.c.
__device__ __forceinline__ void sincos_2pi(float u, float &s, float &c) {
    const float kf = floorf(u * 4.0f + 0.5f);
    const float r = u - kf * 0.25f;                // exact
    const float x = r * 6.28318530717958647692f;   // in [-pi/4, pi/4]
    const float x2 = x * x;
    float sp = fmaf(x2, 2.7557319e-6f, -1.9841270e-4f);
    sp = fmaf(sp, x2, 8.3333333e-3f);
    sp = fmaf(sp, x2, -1.6666667e-1f);
    sp = fmaf(sp * x2, x, x);
    float cp = fmaf(x2, 2.4801587e-5f, -1.3888889e-3f);
    cp = fmaf(cp, x2, 4.1666667e-2f);
    cp = fmaf(cp, x2, -0.5f);
    cp = fmaf(cp, x2, 1.0f);
    const int k = (int)kf & 3;
    s = k == 0 ? sp : k == 1 ? cp : k == 2 ? -sp : -cp;
    c = k == 0 ? cp : k == 1 ? -sp : k == 2 ? -cp : sp;
}

__device__ __forceinline__ float lcg(uint32_t &state) {  // path_tracer.rs:271-279
    state = 1664525u * state + 1013904223u;
    return (float)(state & 0x00ffffffu) * (1.0f / 16777216.0f);
}

struct PtParams {
    AccelView accel;
    const float *const *vertex_heap;   // per instance: [f32; 3] vertices   (vertex_heap.buffer(inst), path_tracer.rs:366)
    const uint32_t *const *index_heap; // per instance: Index triples        (index_heap.buffer(inst),  path_tracer.rs:367)
    float4 *image;                     // Tex2d<Float4> accumulation image, row-major
    uint32_t *seed_image;              // Tex2d<u32>
    uint32_t width, height, spp_per_dispatch, max_depth;
    float tan_half_fov;
    unsigned long long *ray_counters;  // [0] closest-hit rays, [1] any-hit rays traced (for Mrays/s); may be null
};

__global__ void __launch_bounds__(256) k_path_tracer(PtParams P) {
    const uint32_t cx = blockIdx.x * 16 + (threadIdx.x & 15), cy = blockIdx.y * 16 + (threadIdx.x >> 4);
    if (cx >= P.width || cy >= P.height) return;
    const V3 cbox_materials[8] = {{0.725f, 0.710f, 0.680f}, {0.725f, 0.710f, 0.680f}, {0.725f, 0.710f, 0.680f}, {0.140f, 0.450f, 0.091f},
                                  {0.630f, 0.065f, 0.050f}, {0.725f, 0.710f, 0.680f}, {0.725f, 0.710f, 0.680f}, {0.000f, 0.000f, 0.000f}};
    const float FRAC_1_PI = 0.318309886183790671537767526745028724f, F32_MAX = 3.40282347e+38f;
    const float frame_size = (float)min(P.width, P.height);
    uint32_t state = P.seed_image[(size_t)cy * P.width + cx];
    const float rx = lcg(state), ry = lcg(state);
    const float px = ((float)cx + rx) / frame_size * 2.0f - 1.0f, py = ((float)cy + ry) / frame_size * 2.0f - 1.0f;
    V3 radiance = v3(0.f, 0.f, 0.f);
    unsigned long long n_closest = 0, n_any = 0;
    const V3 light_position = v3(-0.24f, 1.98f, 0.16f);
    const V3 light_u = v3(-0.24f, 1.98f, -0.22f) - light_position, light_v = v3(0.23f, 1.98f, 0.16f) - light_position;
    const V3 light_emission = v3(17.0f, 12.0f, 4.0f);
    const float light_area = length(cross(light_u, light_v));
    const V3 light_normal = normalize(cross(light_u, light_v));
    for (uint32_t sample = 0; sample < P.spp_per_dispatch; sample++) {
        // generate_ray(pixel * (1, -1)), path_tracer.rs:291-306
        const V3 cam = v3(-0.01f, 0.995f, 5.0f);
        const V3 pixel = cam + v3(px * 1.0f * P.tan_half_fov, py * -1.0f * P.tan_half_fov, -1.0f);
        V3 ray_o = cam, ray_d = normalize(pixel - cam);
        float ray_tmin = 0.0f, ray_tmax = F32_MAX;
        V3 beta = v3(1.f, 1.f, 1.f);
        float pdf_bsdf = 0.0f;
        uint32_t depth = 0;
        while (depth < P.max_depth) {
            const DeviceHit hit = trace_one<false>(P.accel, make_float4(ray_o.x, ray_o.y, ray_o.z, ray_tmin), make_float4(ray_d.x, ray_d.y, ray_d.z, ray_tmax), 0xffu);
            n_closest++;
            if (hit.inst == kNone) break;
            const float *vb = P.vertex_heap[hit.inst];
            const uint32_t *tri = P.index_heap[hit.inst] + 3 * (size_t)hit.prim;
            const uint32_t i0 = tri[0], i1 = tri[1], i2 = tri[2];
            const V3 p0 = v3(vb[3 * i0], vb[3 * i0 + 1], vb[3 * i0 + 2]), p1 = v3(vb[3 * i1], vb[3 * i1 + 1], vb[3 * i1 + 2]), p2 = v3(vb[3 * i2], vb[3 * i2 + 1], vb[3 * i2 + 2]);
            const V3 p = (p0 * ((1.0f - hit.u) - hit.v) + p1 * hit.u) + p2 * hit.v;  // SurfaceHit::interpolate, rtx.rs:384
            const V3 n = normalize(cross(p1 - p0, p2 - p0));
            const float cos_wi = -dot(ray_d, n);
            if (cos_wi < 1e-4f) break;
            const V3 pp = offset_ray_origin(p, n);
            const V3 albedo = cbox_materials[hit.inst & 7u];
            if (hit.inst == 7u) {  // hit light
                if (depth == 0u) radiance = radiance + light_emission;
                else {
                    const V3 d = p - ray_o;
                    const float pdf_light = dot(d, d) / (light_area * cos_wi);
                    const float mis_weight = pdf_bsdf / fmaxf(pdf_bsdf + pdf_light, 1e-4f);
                    radiance = radiance + (mis_weight * beta) * light_emission;
                }
                break;
            } else {  // sample light
                const float ux_light = lcg(state), uy_light = lcg(state);
                const V3 p_light = (light_position + ux_light * light_u) + uy_light * light_v;
                const V3 pp_light = offset_ray_origin(p_light, light_normal);
                const float d_light = length(pp - pp_light);
                const V3 wi_light = normalize(pp_light - pp);
                const V3 so = offset_ray_origin(pp, n);
                const DeviceHit sh = trace_one<true>(P.accel, make_float4(so.x, so.y, so.z, 0.0f), make_float4(wi_light.x, wi_light.y, wi_light.z, d_light), 0xffu);
                n_any++;
                const bool occluded = sh.inst != kNone;
                const float cos_wi_light = dot(wi_light, n);
                const float cos_light = -dot(light_normal, wi_light);
                if (!occluded && cos_wi_light > 1e-4f && cos_light > 1e-4f) {
                    const float pdf_light = (d_light * d_light) / (light_area * cos_light);
                    const float pdf_b = cos_wi_light * FRAC_1_PI;
                    const float mis_weight = pdf_light / fmaxf(pdf_light + pdf_b, 1e-4f);
                    const V3 bsdf = (albedo * FRAC_1_PI) * cos_wi_light;
                    radiance = radiance + (((beta * bsdf) * mis_weight) * light_emission) / fmaxf(pdf_light, 1e-4f);
                }
            }
            // sample BSDF: make_onb + cosine_sample_hemisphere, path_tracer.rs:311-331
            const V3 binormal = fabsf(n.x) > fabsf(n.z) ? v3(-n.y, n.x, 0.0f) : v3(0.0f, -n.z, n.y);
            const V3 tangent = normalize(cross(binormal, n));
            const float ux = lcg(state), uy = lcg(state);
            const float r = sqrtf(ux);
            float sphi, cphi;
            sincos_2pi(uy, sphi, cphi);
            const V3 local = v3(r * cphi, r * sphi, sqrtf(1.0f - ux));
            const V3 new_direction = (tangent * local.x + binormal * local.y) + n * local.z;
            ray_o = pp; ray_d = new_direction; ray_tmin = 0.0f; ray_tmax = F32_MAX;
            beta = beta * albedo;
            pdf_bsdf = cos_wi * FRAC_1_PI;
            // russian roulette
            const float l = dot(v3(0.212671f, 0.715160f, 0.072169f), beta);
            if (l == 0.0f) break;
            const float q = fmaxf(l, 0.05f);
            const float rr = lcg(state);
            if (rr > q) break;
            beta = beta / q;
            depth += 1;
        }
    }
    radiance = radiance / (float)P.spp_per_dispatch;
    P.seed_image[(size_t)cy * P.width + cx] = state;
    if (isnan(radiance.x) || isnan(radiance.y) || isnan(radiance.z)) radiance = v3(0.f, 0.f, 0.f);
    radiance = v3(fminf(fmaxf(radiance.x, 0.0f), 30.0f), fminf(fmaxf(radiance.y, 0.0f), 30.0f), fminf(fmaxf(radiance.z, 0.0f), 30.0f));
    const float4 old = P.image[(size_t)cy * P.width + cx];
    P.image[(size_t)cy * P.width + cx] = make_float4(radiance.x + old.x, radiance.y + old.y, radiance.z + old.z, old.w + 1.0f);
    if (P.ray_counters) {
        // warp-aggregated by the compiler (REDUX + one atomic per warp)
        atomicAdd(&P.ray_counters[0], n_closest);
        atomicAdd(&P.ray_counters[1], n_any);
    }
}

}  // namespace

void launch_path_tracer(cudaStream_t s, const AccelView &accel, const float *const *vertex_heap, const uint32_t *const *index_heap, float4 *image,
                        uint32_t *seed_image, uint32_t width, uint32_t height, uint32_t spp_per_dispatch, uint32_t max_depth, float tan_half_fov,
                        unsigned long long *ray_counters, LaunchCounter &lc) {
    PtParams P{accel, vertex_heap, index_heap, image, seed_image, width, height, spp_per_dispatch, max_depth, tan_half_fov, ray_counters};
    dim3 grid((width + 15) / 16, (height + 15) / 16);
    k_path_tracer<<<grid, 256, 0, s>>>(P);
    lc.count++;
}

}  // namespace lcb
