"""Curves (CurveBuildCommand api_types:618-631, GeometryImpl::build_curve cpu/accel.rs:142-203, hit encoding
accel.rs:485-501): the oracle's restatement against analytic answers and a float64 sphere-sweep, then the GPU paths (batch entry
points and IR-lowered kernels) against the oracle, bit for bit."""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as ol

LINEAR, BSPLINE, CATMULL, BEZIER = 0, 1, 2, 3
PIECES = 8


# ---- scene helpers ------------------------------------------------------------------------------------------------------------
def strands(basis, n_strands, points_per_strand, seed, step=0.12, radius=(0.01, 0.04)):
    """Random-walk strands in the unit cube: control points (x, y, z, r) and the segments' first-control-point indices."""
    rng = np.random.default_rng(seed)
    cps, segs = [], []
    for _ in range(n_strands):
        base = len(cps)
        p = rng.random(3)
        d = rng.normal(size=3); d /= np.linalg.norm(d)
        for _ in range(points_per_strand):
            cps.append([*p, rng.uniform(*radius)])
            d = d + 0.6 * rng.normal(size=3); d /= np.linalg.norm(d)
            p = p + step * d
        if basis == LINEAR:
            segs += [base + i for i in range(points_per_strand - 1)]
        elif basis == BEZIER:
            segs += [base + i for i in range(0, points_per_strand - 3, 3)]
        else:
            segs += [base + i for i in range(points_per_strand - 3)]
    return np.asarray(cps, np.float32), np.asarray(segs, np.uint32)


def random_rays(n, seed, lo=-0.3, hi=1.3):
    rng = np.random.default_rng(seed)
    rays = np.zeros((n, 8), np.float32)
    rays[:, 0:3] = rng.uniform(lo, hi, (n, 3))
    target = rng.uniform(0.0, 1.0, (n, 3))
    d = target - rays[:, 0:3]
    rays[:, 4:7] = d / np.linalg.norm(d, axis=1, keepdims=True) * rng.uniform(0.5, 2.0, (n, 1))  # unnormalised on purpose
    rays[:, 3] = 1e-4
    rays[:, 7] = 1e30
    return rays


def power_basis(basis, q):
    """float64 copy of the frontend's matrices (lc/src/rtx/curve.rs:88-139): rows a3, a2, a1, a0"""
    q = np.asarray(q, np.float64)
    if basis == BSPLINE:
        m = np.array([[-1, 3, -3, 1], [3, -6, 3, 0], [-3, 0, 3, 0], [1, 4, 1, 0]], np.float64) / 6
    elif basis == CATMULL:
        m = np.array([[-1, 3, -3, 1], [2, -5, 4, -1], [-1, 0, 1, 0], [0, 2, 0, 0]], np.float64) / 2
    else:
        m = np.array([[-1, 3, -3, 1], [3, -6, 3, 0], [-3, 3, 0, 0], [1, 0, 0, 0]], np.float64)
    return m @ q


def sweep_entry(o, d, A, B, samples=4001):
    """float64 reference of one rounded cone: first entry of the ray into the union of the spheres (c(s), r(s)), s on a dense grid"""
    o, d, A, B = (np.asarray(x, np.float64) for x in (o, d, A, B))
    s = np.linspace(0.0, 1.0, samples)[:, None]
    c = A[None, :3] + s * (B[:3] - A[:3])[None, :]
    r = A[3] + s[:, 0] * (B[3] - A[3])
    oc = c - o[None, :]
    dd = d @ d
    b = oc @ d
    disc = b * b - dd * ((oc * oc).sum(1) - r * r)
    ok = disc >= 0
    t = np.where(ok, (b - np.sqrt(np.where(ok, disc, 0))) / dd, np.inf)
    i = int(np.argmin(t))
    return t[i], s[i, 0]


def curve_sweep_entry(o, d, a, samples=40001):
    """float64 reference of one cubic segment: first entry of the ray into the union of the spheres (c(u), r(u)), u on a dense grid of
    [0, 1] (a = power-basis rows a3..a0 as returned by power_basis).  Grid spacing 2.5e-5: the union's surface is within ~1e-9 of the
    envelope for the radii used here."""
    o, d = np.asarray(o, np.float64), np.asarray(d, np.float64)
    u = np.linspace(0.0, 1.0, samples)[:, None]
    c = ((a[0][None, :] * u + a[1][None, :]) * u + a[2][None, :]) * u + a[3][None, :]
    oc = c[:, :3] - o[None, :]
    dd = d @ d
    b = oc @ d
    disc = b * b - dd * ((oc * oc).sum(1) - c[:, 3] ** 2)
    ok = disc >= 0
    t = np.where(ok, (b - np.sqrt(np.where(ok, disc, 0))) / dd, np.inf)
    i = int(np.argmin(t))
    return t[i], u[i, 0], float(disc[i]) / float(dd * max(c[i, 3] ** 2, 1e-300))


# ---- the oracle's restatement (CPU) ---------------------------------------------------------------------------------------------
def test_cone_known_answers():
    # cylinder hit from the side: t = distance - r, parameter = where the ray crosses the axis
    hit, t, s = ol.canonical_cone((0.25, 0, -5), (0, 0, 1), 1e-4, 1e30, (0, 0, 0, 0.1), (1, 0, 0, 0.1))
    assert hit and abs(t - 4.9) < 1e-6 and abs(s - 0.25) < 1e-6
    # end spheres along the axis, unnormalised direction: t scales with 1 / |d|
    hit, t, s = ol.canonical_cone((-5, 0, 0), (1, 0, 0), 1e-4, 1e30, (0, 0, 0, 0.2), (1, 0, 0, 0.1))
    assert hit and abs(t - 4.8) < 1e-6 and s == 0.0
    hit, t, s = ol.canonical_cone((5, 0, 0), (-2, 0, 0), 1e-4, 1e30, (0, 0, 0, 0.2), (1, 0, 0, 0.1))
    assert hit and abs(t - 1.95) < 1e-6 and s == 1.0
    # a miss, a hit beyond tmax, a hit before tmin (origin inside the first sphere: its entry lies behind)
    assert not ol.canonical_cone((0.5, 0.3, -5), (0, 0, 1), 1e-4, 1e30, (0, 0, 0, 0.2), (1, 0, 0, 0.1))[0]
    assert not ol.canonical_cone((0.25, 0, -5), (0, 0, 1), 1e-4, 4.0, (0, 0, 0, 0.1), (1, 0, 0, 0.1))[0]
    # one sphere swallowing the other (no lateral surface): the larger sphere alone
    hit, t, s = ol.canonical_cone((0, 0, -5), (0, 0, 1), 1e-4, 1e30, (0, 0, 0, 0.5), (0.1, 0, 0, 0.1))
    assert hit and abs(t - 4.5) < 1e-6 and s == 0.0


def test_cone_is_the_sphere_sweep():
    """Property: the canonical fp32 answer is the first entry into the union of the swept spheres (float64, dense sampling)."""
    rng = np.random.default_rng(7)
    n_hit = 0
    for _ in range(1500):
        A = np.array([*rng.random(3), rng.uniform(0.02, 0.2)])
        B = np.array([*(A[:3] + rng.normal(size=3) * rng.uniform(0.05, 0.5)), rng.uniform(0.02, 0.2)])
        o = rng.uniform(-3, 4, 3)
        if min(np.linalg.norm(o - A[:3]) - A[3], np.linalg.norm(o - B[:3]) - B[3]) < 0.3:
            continue  # keep the origin outside the solid: the semantic is "entry"
        target = A[:3] + rng.random() * (B[:3] - A[:3]) + rng.normal(size=3) * 0.15
        d = (target - o) * rng.uniform(0.3, 3.0)
        A32, B32, o32, d32 = (x.astype(np.float32) for x in (A, B, o, d))
        hit, t, s = ol.canonical_cone(o32, d32, 0.0, 1e30, A32, B32)
        t_ref, s_ref = sweep_entry(o32, d32, A32, B32)
        if np.isfinite(t_ref):
            # grazing rays may flip; otherwise t agrees to the sampling resolution of the reference
            if hit:
                n_hit += 1
                assert abs(t - t_ref) <= 2e-3 * max(1.0, t_ref), (t, t_ref)
                assert abs(s - s_ref) < 0.05 or abs(t - t_ref) < 1e-4
        else:
            if hit:  # only a graze may differ: the hit point must be within a hair of the surface
                p = o32.astype(np.float64) + t * d32.astype(np.float64)
                c = A32[:3] + s * (B32[:3] - A32[:3]); r = A32[3] + s * (B32[3] - A32[3])
                assert abs(np.linalg.norm(p - c) - r) < 1e-3
    assert n_hit > 300


@pytest.mark.parametrize("basis", [BSPLINE, CATMULL, BEZIER])
def test_basis_matches_frontend_evaluators(basis):
    """A ray aimed at c(u) of the frontend's CubicCurve (rtx/curve.rs) must hit the oracle's curve there: t = distance - r(u) for a
    ray through the centre line, and the reported parameter is u."""
    rng = np.random.default_rng(11 + basis)
    q = np.concatenate([rng.random((4, 3)) * 0.5, rng.uniform(0.02, 0.03, (4, 1))], 1).astype(np.float32)
    o = ol.OracleScene()
    m = o.add_curve(basis, q, [0])
    o.update(1, [{"index": 0, "flags": 1 | 2 | 4 | 16, "visibility": 0xFF, "mesh": m}])
    a = power_basis(basis, q)
    for u in (0.0, 0.125, 0.3, 0.5, 0.77, 1.0):
        c = ((a[0] * u + a[1]) * u + a[2]) * u + a[3]
        tangent = (3 * a[0] * u + 2 * a[1]) * u + a[2]
        n = np.cross(tangent[:3], [0.3, -0.5, 0.8]); n /= np.linalg.norm(n)
        origin = c[:3] + 2.0 * n
        ray = np.zeros((1, 8), np.float32)
        ray[0, 0:3] = origin; ray[0, 3] = 1e-4; ray[0, 4:7] = -n; ray[0, 7] = 1e30
        h = o.trace_closest(ray, mode=ol.BRUTE)[0]
        assert h["inst"] == 0 and h["prim"] == 0
        assert h["bary"][1] == -1.0  # the curve marker (cpu/accel.rs:491-494)
        # the hit is on the true swept surface (the cone pieces only locate it, refine_curve_hit): equal to the dense float64 sweep, which
        # enters no later than the sphere the ray was aimed at
        t_ref, u_ref, _ = curve_sweep_entry(ray[0, 0:3], ray[0, 4:7], a)
        assert t_ref <= 2.0 - c[3] + 1e-6
        assert abs(h["committed_ray_t"] - t_ref) <= 1e-5 * t_ref, (u, h["committed_ray_t"], t_ref)
        assert abs(h["bary"][0] - u_ref) < 2e-3 and abs(h["bary"][0] - u) < 0.05
    o.close()


def oracle_curve_scene(basis, seed, with_mesh=True, opaque=True):
    cps, segs = strands(basis, 40, 8 if basis != BEZIER else 10, seed)
    o = ol.OracleScene()
    cm = o.add_curve(basis, cps, segs)
    affine = np.array([[0.9, 0.1, 0, 0.05], [-0.1, 0.9, 0, 0.02], [0, 0, 1.1, -0.03]], np.float32)
    mods = [{"index": 0, "flags": 1 | 2 | (4 if opaque else 8) | 16, "visibility": 0xFF, "mesh": cm, "affine": affine}]
    quad = None
    if with_mesh:
        verts = np.array([[-0.5, -0.5, 0.5], [1.5, -0.5, 0.5], [1.5, 1.5, 0.5], [-0.5, 1.5, 0.5]], np.float32)
        tris = np.array([[0, 1, 2], [0, 2, 3]], np.uint32)
        tm = o.add_mesh(verts, tris)
        mods.append({"index": 1, "flags": 1 | 2 | 4 | 16, "visibility": 0x0F, "mesh": tm})
        quad = (verts, tris)
    o.update(len(mods), mods)
    return o, cps, segs, affine, quad


@pytest.mark.parametrize("basis", [BSPLINE, CATMULL, BEZIER])
def test_curve_hits_agree_with_the_float64_swept_sphere_reference(basis):
    """north_star's tolerance on curves: t within 1e-5 relative of a dense float64 sweep of the sphere (c(u), r(u)) along the true cubic —
    not along the cone pieces.  Rays that graze the surface (the entry sphere is cut at less than 2 % of its radius squared) are where a
    locate-then-refine intersector and the sweep may legitimately disagree about hit / miss, and are skipped; everywhere else: same
    hit / miss, same t, same u."""
    rng = np.random.default_rng(40 + basis)
    n_checked = 0
    for trial in range(10):
        q = np.concatenate([np.cumsum(rng.normal(size=(4, 3)) * 0.25, 0), rng.uniform(0.02, 0.08, (4, 1))], 1).astype(np.float32)
        o = ol.OracleScene()
        m = o.add_curve(basis, q, [0])
        o.update(1, [{"index": 0, "flags": 1 | 2 | 4 | 16, "visibility": 0xFF, "mesh": m}])
        a = power_basis(basis, q)
        n = 50
        rays = np.zeros((n, 8), np.float32)
        for i in range(n):
            u = rng.random()
            c = ((a[0] * u + a[1]) * u + a[2]) * u + a[3]
            origin = c[:3] + rng.normal(size=3) * 1.5
            target = c[:3] + rng.normal(size=3) * c[3] * 0.6
            d = (target - origin) * rng.uniform(0.4, 2.5)       # unnormalised on purpose
            rays[i, 0:3] = origin; rays[i, 3] = 1e-4; rays[i, 4:7] = d; rays[i, 7] = 1e30
        hits = o.trace_closest(rays, mode=ol.BRUTE)
        for i in range(n):
            t_ref, u_ref, margin = curve_sweep_entry(rays[i, 0:3], rays[i, 4:7], a)
            if not np.isfinite(t_ref) or margin < 0.02 or t_ref <= 1e-3:
                continue
            assert hits["inst"][i] == 0, (trial, i, t_ref)
            assert abs(hits["committed_ray_t"][i] - t_ref) <= 1e-5 * abs(t_ref), (trial, i, hits["committed_ray_t"][i], t_ref)
            if 1e-3 < u_ref < 1 - 1e-3:
                assert abs(hits["bary"][i, 0] - u_ref) < 1e-3, (trial, i, hits["bary"][i, 0], u_ref)
            n_checked += 1
        o.close()
    assert n_checked > 200


@pytest.mark.parametrize("basis", [LINEAR, BSPLINE, CATMULL, BEZIER])
def test_oracle_curve_scene_consistency(basis):
    o, cps, segs, _, _ = oracle_curve_scene(basis, 100 + basis)
    rays = random_rays(4000, 5 + basis)
    hits = o.trace_closest(rays, mode=ol.BRUTE)
    occ = o.trace_any(rays, mode=ol.BRUTE)
    assert np.array_equal(occ != 0, hits["inst"] != 0xFFFFFFFF)
    on_curve = hits["inst"] == 0
    assert on_curve.sum() > 100 and (hits["inst"] == 1).sum() > 100
    assert np.all(hits["bary"][on_curve, 1] == -1.0) and np.all(hits["prim"][on_curve] < segs.shape[0])
    assert np.all((hits["bary"][on_curve, 0] >= 0) & (hits["bary"][on_curve, 0] <= 1))
    # masks: the quad is invisible to mask 0xF0, the curve is not
    hits2 = o.trace_closest(rays, mask=0xF0, mode=ol.BRUTE)
    assert not np.any(hits2["inst"] == 1) and (hits2["inst"] == 0).sum() >= on_curve.sum()
    # every reported point lies on the sphere of the sweep the parameter names (object space)
    o.close()


# ---- GPU ------------------------------------------------------------------------------------------------------------------------
def device_curve_scene(device, lc, basis, cps, segs, affine, quad, opaque=True, cp_stride_floats=4):
    keep = []
    if cp_stride_floats != 4:
        padded = np.zeros((cps.shape[0], cp_stride_floats), np.float32); padded[:, :4] = cps
        cpb = device.create_buffer_from_array(padded)
    else:
        cpb = device.create_buffer_from_array(cps)
    sgb = device.create_buffer_from_array(segs)
    curve = device.create_curve(basis, cpb.view(), sgb.view())
    curve.build()
    accel = device.create_accel()
    t = np.eye(4, dtype=np.float32); t[:3, :] = affine
    accel.push_curve(curve, t, 0xFF, opaque)
    keep += [cpb, sgb, curve]
    if quad is not None:
        vb, ib = device.create_buffer_from_array(quad[0]), device.create_buffer_from_array(quad[1])
        mesh = device.create_mesh(vb.view(), ib.view())
        mesh.build()
        accel.push_mesh(mesh, None, 0x0F, True)
        keep += [vb, ib, mesh]
    accel.build()
    return accel, keep


def batch_closest(device, lc, accel, rays, mask=0xFF):
    n = rays.shape[0]
    rb, hb = device.create_buffer(n, 32, 16), device.create_buffer(n, 24, 8)
    rb.view(0, n).copy_from(rays)
    accel.intersect(rb.view(0, n), hb.view(0, n), n, mask)
    hits = np.zeros(n, dtype=lc.SurfaceHit)
    hb.view(0, n).copy_to(hits)
    rb.destroy(); hb.destroy()
    return hits


@pytest.mark.gpu
@pytest.mark.parametrize("basis", [LINEAR, BSPLINE, CATMULL, BEZIER])
def test_gpu_curves_match_oracle(device, basis):
    import luisa_compute_rs_b200 as lc
    from harness import assert_hits_equal
    o, cps, segs, affine, quad = oracle_curve_scene(basis, 100 + basis)
    accel, keep = device_curve_scene(device, lc, basis, cps, segs, affine, quad, cp_stride_floats=4 if basis != CATMULL else 8)
    rays = random_rays(60000, 21 + basis)
    want = o.trace_closest(rays, mode=ol.BRUTE)
    got = batch_closest(device, lc, accel, rays)
    assert (want["inst"] == 0).sum() > 2000
    assert_hits_equal(got, want, f"curves basis {basis}")
    # masked, and any-hit
    assert_hits_equal(batch_closest(device, lc, accel, rays, 0xF0), o.trace_closest(rays, mask=0xF0, mode=ol.BRUTE), "curves, mask 0xF0")
    n = rays.shape[0]
    rb, ob = device.create_buffer(n, 32, 16), device.create_buffer(n, 4, 4)
    rb.view(0, n).copy_from(rays)
    accel.intersect_any(rb.view(0, n), ob.view(0, n), n, 0xFF)
    occ = np.zeros(n, np.uint32); ob.view(0, n).copy_to(occ)
    assert np.array_equal(occ, o.trace_any(rays, mode=ol.BRUTE))
    o.close()


@pytest.mark.gpu
def test_gpu_curve_ray_query(device):
    """Non-opaque curve instance: candidates go through the surface-candidate hook with bary = (u, -1) (cpu/accel.rs:650-684)."""
    import luisa_compute_rs_b200 as lc
    o, cps, segs, affine, quad = oracle_curve_scene(BSPLINE, 301, opaque=False)
    accel, keep = device_curve_scene(device, lc, BSPLINE, cps, segs, affine, quad, opaque=False)
    rays = random_rays(30000, 77)
    n = rays.shape[0]
    rb, hb = device.create_buffer(n, 32, 16), device.create_buffer(n, 24, 8)
    rb.view(0, n).copy_from(rays)
    for kind, radius in ((0, 0.0), (3, 0.0), (1, 1.2)):
        accel.traverse(rb.view(0, n), hb.view(0, n), n, 0xFF, lc.SurfaceCandidateFilter(kind, radius))
        device.default_stream().synchronize()
        got = np.zeros(n, dtype=lc.CommittedHit); hb.view(0, n).copy_to(got)
        want = o.ray_query(rays, kind=kind, radius=radius, mode=ol.BRUTE)
        for f in ("inst", "prim", "hit_type"):
            assert np.array_equal(got[f], want[f]), (kind, f)
        assert np.array_equal(got["bary"].view(np.uint32), want["bary"].view(np.uint32))
        assert np.array_equal(got["committed_ray_t"].view(np.uint32), want["committed_ray_t"].view(np.uint32))
        if kind == 3:
            assert not np.any(got["inst"] == 0)  # every curve candidate rejected: only the opaque quad is left
    o.close()


def closest_kernel(lc, curve_bases):
    """one ray per thread: hits[i] = accel.intersect(rays[i], mask) with AccelTraceOptions::curve_bases recorded in the module"""
    from luisa_compute_rs_b200 import ir
    from luisa_compute_rs_b200.examples_ir import common_types
    k = ir.KernelBuilder(block_size=(128, 1, 1), curve_bases=curve_bases)
    f3, ray_ty, hit_ty = common_types(k)
    rays, out, accel = k.arg_buffer(ray_ty), k.arg_buffer(hit_ty), k.arg_accel()

    def body():
        i = k.dispatch_id().x
        out.write(i, accel.trace_closest(rays.read(i), 0xFF, hit_ty))
    k.body(body)
    k.finish()
    return k


@pytest.mark.gpu
def test_lowered_kernel_traces_curves(device):
    """The same accel from inside an IR-lowered kernel: with a curve basis recorded the hits equal the oracle's; without one the curve
    instance is skipped (what the OptiX backend's pipeline would do) and only the quad is seen."""
    import luisa_compute_rs_b200 as lc
    from harness import assert_hits_equal
    o, cps, segs, affine, quad = oracle_curve_scene(CATMULL, 404)
    accel, keep = device_curve_scene(device, lc, CATMULL, cps, segs, affine, quad)
    rays = random_rays(40000, 91)
    n = rays.shape[0]
    rb, hb = device.create_buffer(n, 32, 16), device.create_buffer(n, 24, 8)
    rb.view(0, n).copy_from(rays)
    stream = device.default_stream()
    for bases in (4, 0):
        k = closest_kernel(lc, bases)
        shader = device.create_shader(C.addressof(k.km), keep=k)
        stream.submit([shader.dispatch_async((n,), rb, hb, accel)])
        stream.synchronize()
        got = np.zeros(n, dtype=lc.SurfaceHit); hb.view(0, n).copy_to(got)
        if bases:
            assert_hits_equal(got, o.trace_closest(rays, mode=ol.BRUTE), "lowered kernel, curves on")
        else:
            o2 = ol.OracleScene()
            tm = o2.add_mesh(*quad)
            o2.update(2, [{"index": 1, "flags": 1 | 2 | 4 | 16, "visibility": 0x0F, "mesh": tm}])
            assert_hits_equal(got, o2.trace_closest(rays, mode=ol.BRUTE), "lowered kernel, curves off")
            o2.close()
        shader.destroy()
    o.close()


def test_curve_kernel_compiles_with_curve_support():
    """CPU: the lowering defines LCB_CURVES exactly when the module records a curve basis, and the kernel compiles either way."""
    import luisa_compute_rs_b200 as lc
    L = lc._abi.load_library()
    for bases in (0, 2):
        k = closest_kernel(lc, bases)
        text = C.string_at(L.lc_b200_ir_lower_source(C.addressof(k.km))).decode()
        assert ("#define LCB_CURVES 1" in text) == bool(bases)
        log = C.c_void_p()
        assert L.lc_b200_shader_compile_check(C.addressof(k.km), False, C.byref(log)) == 0, C.string_at(log).decode()


@pytest.mark.gpu
def test_curve_rebuild_follows_the_control_points(device):
    """CurveBuild reads the control points where they lie (shared buffers, cpu/accel.rs:160-181): after the buffer changes, a
    rebuild (PreferUpdate is served by a full build) and an AccelBuild give the oracle's hits for the new curve."""
    import luisa_compute_rs_b200 as lc
    from harness import assert_hits_equal
    cps, segs = strands(BEZIER, 30, 10, 900)
    cpb, sgb = device.create_buffer_from_array(cps), device.create_buffer_from_array(segs)
    curve = device.create_curve(lc.CurveBasis.BEZIER, cpb.view(), sgb.view(), lc.AccelOption(allow_update=True))
    curve.build()
    accel = device.create_accel()
    accel.push_curve(curve)
    accel.build()
    rays = random_rays(20000, 901)
    for step in range(2):
        o = ol.OracleScene()
        m = o.add_curve(BEZIER, cps, segs)
        o.update(1, [{"index": 0, "flags": 1 | 2 | 4 | 16, "visibility": 0xFF, "mesh": m}])
        assert_hits_equal(batch_closest(device, lc, accel, rays), o.trace_closest(rays, mode=ol.BRUTE), f"curve rebuild step {step}")
        o.close()
        cps = cps.copy(); cps[:, :3] += np.float32(0.05) * np.sin(np.arange(cps.shape[0], dtype=np.float32))[:, None]; cps[:, 3] *= np.float32(1.5)
        cpb.view().copy_from(cps)
        curve.build(lc.AccelBuildRequest.PREFER_UPDATE)
        accel.build()
