"""GPU parity tests: every query goes through the C ABI (luisa_compute_lib_interface -> DeviceInterface.dispatch and the
lc_b200_* batch entry points) and is compared bit-for-bit with the oracle on the same inputs."""
import os

import numpy as np
import pytest

import luisa_compute_rs_b200 as lc
import oracle_lib as ol
import scenes
from harness import DeviceScene, assert_hits_equal, compare_with_truth

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def check_scene(device, desc, rays, masks=(0xFF,), stride=12, any_hit=True):
    o = ol.scene_from_desc(desc)
    d = DeviceScene(device, desc, vertex_stride=stride)
    try:
        for mask in masks:
            want = o.trace_closest(rays, mask, ol.BVH)
            got = d.trace_closest(rays, mask)
            assert_hits_equal(got, want, f"mask {mask:#x}")
            if any_hit:
                assert np.array_equal(d.trace_any(rays, mask), o.trace_any(rays, mask, ol.BVH))
    finally:
        d.destroy(); o.close()


def test_device_identity(device):
    assert device.name() == "b200"
    assert device.query("no_such_property") is None


def test_c1_triangle(device):
    rays = scenes.c1_rays(1024, 1024)  # C1 at BASELINE size
    check_scene(device, scenes.c1_triangle(), rays)
    g = np.load(os.path.join(GOLDEN, "c1_closed_form_64.npz"))
    d = DeviceScene(device, scenes.c1_triangle())
    hits = d.trace_closest(scenes.c1_rays(64, 64))
    clear = g["edge_margin"] > 1e-5
    assert np.array_equal((hits["inst"] != lc.INVALID)[clear], g["hit"][clear])
    v = (hits["inst"] != lc.INVALID) & clear
    assert np.max(np.abs(hits["committed_ray_t"][v] - g["t"][v]) / g["t"][v]) < 1e-5
    assert np.max(np.abs(hits["bary"][v] - g["bary"][v])) < 1e-5
    d.destroy()


def test_c2_cornell_primary_and_golden(device):
    desc = scenes.c2_cornell()
    check_scene(device, desc, scenes.c2_primary_rays(1024, 1024))  # C2 primary rays at BASELINE size
    g = np.load(os.path.join(GOLDEN, "cornell_primary_48.npz"))
    d = DeviceScene(device, desc)
    hits = d.trace_closest(scenes.c2_primary_rays(48, 48))
    assert np.array_equal(hits["inst"], g["inst"]) and np.array_equal(hits["prim"], g["prim"])
    assert np.array_equal(hits["committed_ray_t"].view(np.uint32), g["t_bits"])
    assert np.array_equal(hits["bary"].view(np.uint32), g["bary_bits"])
    d.destroy()


def test_c2_cornell_secondary_rays(device):
    """diffuse bounce + shadow rays started on surfaces with offset_ray_origin (path_tracer.rs:383-428)"""
    desc = scenes.c2_cornell()
    o = ol.scene_from_desc(desc)
    prim = scenes.c2_primary_rays(128, 128)
    h = o.trace_closest(prim)
    v = h["inst"] != lc.INVALID
    p = prim["orig"][v] + prim["dir"][v] * h["committed_ray_t"][v][:, None]
    rng = np.random.default_rng(5)
    d = rng.normal(size=p.shape).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    sec = scenes.make_rays(lc.offset_ray_origin(p, d), d, 0.0, np.float32(3.4e38))
    dev = DeviceScene(device, desc)
    assert_hits_equal(dev.trace_closest(sec), o.trace_closest(sec), "secondary")
    light = np.array([-0.005, 1.98, -0.03], np.float32)
    dl = light - p
    dist = np.linalg.norm(dl, axis=1).astype(np.float32)
    sh = scenes.make_rays(p, dl / dist[:, None], 1e-4, dist * np.float32(0.999))
    assert np.array_equal(dev.trace_any(sh), o.trace_any(sh))
    dev.destroy(); o.close()


@pytest.mark.parametrize("n_tris,n_rays,seed", [(1, 2000, 1), (2, 2000, 2), (3, 2000, 3), (4, 3000, 4), (9, 3000, 5), (100, 20000, 6),
                                                (5000, 100000, 7), (200000, 400000, 8)])
def test_soup_closest_and_any(device, n_tris, n_rays, seed):
    desc = scenes.c3_soup(n_tris, seed=seed)
    check_scene(device, desc, scenes.incoherent_rays(n_rays, seed=100 + seed))


def test_soup_vs_f64_truth_and_tie_rate(device):
    desc = scenes.c3_soup(50000)
    rays = scenes.incoherent_rays(300000)
    o = ol.scene_from_desc(desc)
    d = DeviceScene(device, desc)
    got = d.trace_closest(rays)
    truth, amb = o.truth(rays)
    r = compare_with_truth(got, truth, amb)
    print("truth comparison:", r)
    assert r["mismatches"] == 0           # inst/prim identical outside reported ties
    assert r["t_rel_err"] < 1e-5 and r["bary_abs_err"] < 1e-5
    assert r["tie_rate"] < 0.01
    d.destroy(); o.close()


def test_float3_stride_16_vertices(device):
    check_scene(device, scenes.c3_soup(3000, seed=21), scenes.incoherent_rays(20000, seed=22), stride=16)


def test_instances_transforms_masks_user_ids(device):
    desc = scenes.instanced_scene(2000, 10)
    rays = scenes.incoherent_rays(200000, lo=-1.0, hi=8.0, seed=31)
    check_scene(device, desc, rays, masks=(0xFF, 0xF0, 0x01, 0x0))
    d = DeviceScene(device, desc)
    o = ol.scene_from_desc(desc)
    for i in range(10):
        assert d.accel.instance_user_id(i) == o.instance_user_id(i) == 100 + i
        assert d.accel.instance_visibility_mask(i) == o.instance_visibility(i)
        assert np.array_equal(d.accel.instance_transform(i), o.instance_transform(i))
    d.destroy(); o.close()


def test_scaled_and_sheared_instance(device):
    s = scenes.SceneDesc()
    m = s.add_mesh(*scenes.random_soup(500, 41, extent=0.1))
    s.add_instance(m, np.array([[2.0, 0.3, 0, 1], [0, 0.5, 0.1, -2], [0.2, 0, 3.0, 0.5]], np.float32))
    s.add_instance(m, np.array([[0.1, 0, 0, 3], [0, 0.1, 0, 0], [0, 0, 0.1, 0]], np.float32))
    check_scene(device, s, scenes.incoherent_rays(100000, lo=-3.0, hi=5.0, seed=42))


def test_ties_lowest_ids_and_coincident_geometry(device):
    s = scenes.SceneDesc()
    v = [[0, 0, 1], [1, 0, 1], [0, 1, 1]]
    m = s.add_mesh(v + v + v, [[6, 7, 8], [3, 4, 5], [0, 1, 2]])
    s.add_instance(m); s.add_instance(m); s.add_instance(m)
    rng = np.random.default_rng(9)
    o = np.concatenate([rng.random((5000, 2), dtype=np.float32), np.zeros((5000, 1), np.float32)], 1)
    rays = scenes.make_rays(o, np.tile(np.float32([0, 0, 1]), (5000, 1)), 0.0, 10.0)
    check_scene(device, s, rays)
    d = DeviceScene(device, s)
    h = d.trace_closest(rays)
    hit = h["inst"] != lc.INVALID
    assert hit.sum() > 1000 and np.all(h["inst"][hit] == 0) and np.all(h["prim"][hit] == 0)
    d.destroy()


def test_shared_edges_are_watertight_grid(device):
    verts, tris = scenes.terrain(65)
    s = scenes.SceneDesc(); s.add_instance(s.add_mesh(verts, tris))
    rng = np.random.default_rng(10)
    n = 200000
    o = np.stack([rng.random(n, dtype=np.float32) * 0.98 + 0.01, np.full(n, 2.0, np.float32), rng.random(n, dtype=np.float32) * 0.98 + 0.01], 1)
    # a share of rays aimed exactly at grid vertices and edges
    gv = verts[rng.integers(0, verts.shape[0], n // 4)]
    o[: n // 4, 0] = np.clip(gv[:, 0], 0.01, 0.99); o[: n // 4, 2] = np.clip(gv[:, 2], 0.01, 0.99)
    rays = scenes.make_rays(o, np.tile(np.float32([0, -1, 0]), (n, 1)), 0.0, 10.0)
    check_scene(device, s, rays)
    d = DeviceScene(device, s)
    assert (d.trace_closest(rays)["inst"] != lc.INVALID).all()   # no ray slips between triangles
    d.destroy()


def test_terrain_grazing_and_axis_aligned_rays(device):
    verts, tris = scenes.terrain(200)
    s = scenes.SceneDesc(); s.add_instance(s.add_mesh(verts, tris))
    rng = np.random.default_rng(12)
    n = 100000
    o = np.stack([rng.random(n, dtype=np.float32), rng.random(n, dtype=np.float32) * 0.2, rng.random(n, dtype=np.float32)], 1)
    d = np.zeros((n, 3), np.float32)
    ax = rng.integers(0, 3, n)
    d[np.arange(n), ax] = rng.choice(np.float32([-1, 1]), n)     # exactly axis-aligned: zero direction components
    d[n // 2:] += rng.normal(scale=1e-3, size=(n - n // 2, 3)).astype(np.float32)
    check_scene(device, s, scenes.make_rays(o, d, 0.0, 10.0))


def test_hit_interval_and_backfaces(device):
    s = scenes.SceneDesc()
    s.add_instance(s.add_mesh([[0, 0, 1], [1, 0, 1], [0, 1, 1]], [[0, 1, 2]]))
    o = ol.scene_from_desc(s)
    t = o.trace_closest(scenes.make_rays([[0.2, 0.2, 0]], [[0, 0, 1]], 0.0, 10.0))["committed_ray_t"][0]
    below = np.nextafter(t, np.float32(0))
    rays = scenes.make_rays([[0.2, 0.2, 0]] * 5 + [[0.2, 0.2, 2]], [[0, 0, 1]] * 5 + [[0, 0, -1]], [0, t, below, 0, 0, 0], [t, 2, 2, below, 10, 10])
    check_scene(device, s, rays)
    d = DeviceScene(device, s)
    assert (d.trace_closest(rays)["inst"] != lc.INVALID).tolist() == [True, False, True, False, True, True]
    d.destroy()


def test_empty_inputs(device):
    s = scenes.SceneDesc()
    s.add_instance(s.add_mesh(*scenes.random_soup(50, 3)))
    d = DeviceScene(device, s)
    assert d.trace_closest(np.zeros(0, dtype=lc.Ray)).shape == (0,)
    d.destroy()
    # accel without instances, and an instance of an empty mesh: every ray misses with t = tmax
    rays = scenes.incoherent_rays(1000)
    empty = scenes.SceneDesc()
    d = DeviceScene(device, empty)
    h = d.trace_closest(rays)
    assert (h["inst"] == lc.INVALID).all() and np.array_equal(h["committed_ray_t"], rays["tmax"]) and not d.trace_any(rays).any()
    d.destroy()
    e2 = scenes.SceneDesc()
    e2.add_instance(e2.add_mesh(np.zeros((0, 3), np.float32), np.zeros((0, 3), np.uint32)))
    e2.add_instance(e2.add_mesh(*scenes.random_soup(10, 4)))
    check_scene(device, e2, rays)


def test_degenerate_triangles_and_zero_direction(device):
    s = scenes.SceneDesc()
    s.add_instance(s.add_mesh([[0, 0, 1], [0, 0, 1], [0, 0, 1], [0, 0, 2], [1, 0, 2], [2, 0, 2], [0, 0, 3], [1, 0, 3], [0, 1, 3]], [[0, 1, 2], [3, 4, 5], [6, 7, 8]]))
    rays = scenes.make_rays([[0.1, 0.1, 0], [0, 0, 0], [0.5, 0, 0], [0, 0, 0]], [[0, 0, 1], [0, 0, 0], [0, 0, 1], [0, 0, 1]], 0.0, 10.0)
    check_scene(device, s, rays)


def test_duplicate_centroids_deep_tree(device):
    """thousands of triangles with identical Morton codes exercise the equal-key split path"""
    rng = np.random.default_rng(13)
    n = 6000
    base = np.float32([[0, 0, 0], [1, 0, 0], [0, 1, 0]])
    v = np.tile(base, (n, 1, 1)).astype(np.float32)
    v[:, :, 2] = (rng.integers(0, 4, n) * np.float32(0.25))[:, None]   # only 4 distinct planes
    s = scenes.SceneDesc()
    s.add_instance(s.add_mesh(v.reshape(-1, 3), np.arange(3 * n, dtype=np.uint32).reshape(-1, 3)))
    o = np.concatenate([rng.random((20000, 2), dtype=np.float32) * 0.5, np.full((20000, 1), -1, np.float32)], 1)
    check_scene(device, s, scenes.make_rays(o, np.tile(np.float32([0, 0, 1]), (20000, 1)), 0.0, 10.0))


def test_rebuild_after_vertex_update_and_accel_modifications(device):
    verts, tris = scenes.random_soup(4000, 51, extent=0.05)
    rays = scenes.incoherent_rays(50000, seed=52)
    vb = device.create_buffer_from_array(verts); ib = device.create_buffer_from_array(tris)
    mesh = device.create_mesh(vb.view(), ib.view(), lc.AccelOption(allow_update=True))
    mesh.build()
    accel = device.create_accel()
    accel.push_mesh(mesh); accel.build()
    o = ol.OracleScene()
    ov = verts.copy()
    om = o.add_mesh(ov, tris)
    o.update(1, [dict(index=0, flags=1 | 2 | 4 | 16 | 32, visibility=0xFF, mesh=om)])
    assert_hits_equal(accel.intersect_host(rays), o.trace_closest(rays), "initial")
    # the BVH aliases user buffers (accel.rs:217-237): rewrite vertices, PreferUpdate, rebuild the accel
    ov += np.float32(0.01) * np.sin(40 * ov[:, [1, 2, 0]]).astype(np.float32)
    vb.view().copy_from(ov); o.commit_mesh(om)
    mesh.build(lc.AccelBuildRequest.PREFER_UPDATE); accel.build(lc.AccelBuildRequest.PREFER_UPDATE)
    assert_hits_equal(accel.intersect_host(rays), o.trace_closest(rays), "after vertex update")
    # second instance, moved; then visibility change; then pop
    t = np.eye(4, dtype=np.float32); t[:3, 3] = [0.5, 0.1, -0.2]
    accel.push_mesh(mesh, t, 0x0F); accel.build()
    o.update(2, [dict(index=1, flags=1 | 2 | 4 | 16 | 32, visibility=0x0F, mesh=om, affine=t[:3].reshape(12))])
    for mask in (0xFF, 0xF0):
        assert_hits_equal(accel.intersect_host(rays, mask), o.trace_closest(rays, mask), f"two instances {mask:#x}")
    accel.set_visibility_on_update(0, 0x10); accel.set_transform_on_update(1, np.eye(4, dtype=np.float32)); accel.build()
    o.update(2, [dict(index=0, flags=16, visibility=0x10), dict(index=1, flags=2)])
    for mask in (0xFF, 0x10, 0x0F):
        assert_hits_equal(accel.intersect_host(rays, mask), o.trace_closest(rays, mask), f"modified {mask:#x}")
    accel.pop(); accel.build(); o.update(1, [])
    assert_hits_equal(accel.intersect_host(rays), o.trace_closest(rays), "after pop")
    assert np.array_equal(accel.intersect_any_host(rays), o.trace_any(rays))
    accel.destroy(); mesh.destroy(); vb.destroy(); ib.destroy(); o.close()


def test_prefer_update_refit_terrain_frames(device):
    """C4-shaped: a height field animated over frames; MeshBuild(PreferUpdate) on an updatable mesh refits in place
    (stats.was_refit), a non-updatable mesh rebuilds; both must give the oracle's hits for the CURRENT vertices."""
    nx = 120
    verts, tris = scenes.terrain(nx)
    rng = np.random.default_rng(77)
    n = 60000
    o = np.stack([rng.random(n, dtype=np.float32), np.full(n, 0.5, np.float32), rng.random(n, dtype=np.float32)], 1)
    d = rng.normal(size=(n, 3)).astype(np.float32); d[:, 1] = -np.abs(d[:, 1]) - 0.2
    rays = scenes.make_rays(o, d, 0.0, 10.0)
    ora = ol.OracleScene()
    ov = verts.copy()
    om = ora.add_mesh(ov, tris)
    ora.update(1, [dict(index=0, flags=1 | 2 | 4 | 16 | 32, visibility=0xFF, mesh=om)])
    objs = []
    for allow in (True, False):
        vb = device.create_buffer_from_array(verts); ib = device.create_buffer_from_array(tris)
        mesh = device.create_mesh(vb.view(), ib.view(), lc.AccelOption(allow_update=allow))
        mesh.build()
        accel = device.create_accel(); accel.push_mesh(mesh); accel.build()
        objs.append((vb, ib, mesh, accel))
    for frame in range(1, 5):
        fv, _ = scenes.terrain(nx, frame=frame * 7)
        if frame == 3:
            fv = fv.copy(); fv[:, 1] += np.float32(0.3)      # large displacement: boxes must follow, not just grow
        ov[:] = fv; ora.commit_mesh(om)
        want = ora.trace_closest(rays)
        for (vb, ib, mesh, accel), allow in zip(objs, (True, False)):
            vb.view().copy_from(fv)
            mesh.build(lc.AccelBuildRequest.PREFER_UPDATE); accel.build(lc.AccelBuildRequest.PREFER_UPDATE)
            assert mesh.stats()["was_refit"] == (1 if allow else 0)
            assert_hits_equal(accel.intersect_host(rays), want, f"frame {frame} allow_update={allow}")
            assert np.array_equal(accel.intersect_any_host(rays), ora.trace_any(rays))
    # a ForceBuild after refits starts from scratch again
    vb, ib, mesh, accel = objs[0]
    mesh.build(lc.AccelBuildRequest.FORCE_BUILD); accel.build()
    assert mesh.stats()["was_refit"] == 0
    assert_hits_equal(accel.intersect_host(rays), ora.trace_closest(rays), "rebuild after refits")
    for vb, ib, mesh, accel in objs:
        accel.destroy(); mesh.destroy(); vb.destroy(); ib.destroy()
    ora.close()


def test_ray_query_candidates_and_opaque_instances(device):
    """Batch RayQuery (traverse / traverse_any) against the oracle: mixed opaque / non-opaque instances, every candidate hook."""
    s = scenes.SceneDesc()
    m0 = s.add_mesh(*scenes.random_soup(3000, 71, extent=0.08))
    m1 = s.add_mesh(*scenes.random_soup(500, 72, extent=0.15))
    s.add_instance(m0, opaque=False)
    t = scenes.rotation_y(30.0); t[:, 3] = [0.2, 0.1, -0.1]
    s.add_instance(m1, t, opaque=True)
    t2 = scenes.IDENTITY34.copy(); t2[:, 3] = [-0.1, 0.0, 0.2]
    s.add_instance(m1, t2, opaque=False, mask=0x0F)
    rays = scenes.incoherent_rays(150000, seed=73)
    o = ol.scene_from_desc(s)
    d = DeviceScene(device, s)
    n = rays.shape[0]
    rb = device.create_buffer_from_array(rays); hb = device.create_buffer(n, 24, 8)
    rng = np.random.default_rng(74)
    first_bit = np.array([0, 3000, 3500], np.uint32)
    bits = rng.integers(0, 2**32, 4000 // 32 + 1, dtype=np.uint64).astype(np.uint32)
    bb = device.create_buffer_from_array(bits); fb = device.create_buffer_from_array(first_bit)
    cases = [(lc.SurfaceCandidateFilter.commit_all(), dict(kind=0)), (lc.SurfaceCandidateFilter.reject_all(), dict(kind=3)),
             (lc.SurfaceCandidateFilter.bary_disc(0.8), dict(kind=1, radius=0.8)), (lc.SurfaceCandidateFilter.bary_disc(0.55), dict(kind=1, radius=0.55)),
             (lc.SurfaceCandidateFilter.prim_bits(bb, fb), dict(kind=2, bits=bits, first_bit=first_bit))]
    for flt, kw in cases:
        for mask in (0xFF, 0xF0):
            d.accel.traverse(rb, hb, n, mask, flt)
            got = hb.view().to_numpy(lc.CommittedHit)
            want = o.ray_query(rays, mask, False, **kw)
            assert got.tobytes() == want.tobytes(), f"traverse {kw.get('kind')} mask {mask:#x}: {(got != want).sum()} rays differ"
            d.accel.traverse_any(rb, hb, n, mask, flt)
            got_any = hb.view().to_numpy(lc.CommittedHit)
            assert np.array_equal(got_any["hit_type"], want["hit_type"])
            h = got_any["hit_type"] == 1       # the reported first hit is a real committed candidate inside the ray interval
            assert np.all(got_any["committed_ray_t"][h] > rays["tmin"][h]) and np.all(got_any["committed_ray_t"][h] >= want["committed_ray_t"][h])
    # commit-all equals trace_closest
    d.accel.traverse(rb, hb, n, 0xFF, None)
    q = hb.view().to_numpy(lc.CommittedHit)
    c = d.trace_closest(rays)
    assert np.array_equal(q["inst"], c["inst"]) and np.array_equal(q["prim"], c["prim"]) and np.array_equal(q["bary"], c["bary"])
    for b in (rb, hb, bb, fb):
        b.destroy()
    d.destroy(); o.close()


def test_c2_path_tracer_image_is_bit_identical_to_the_cpu_restatement(device):
    """Config C2: the kernel of examples/path_tracer.rs (hand-lowered, csrc/path_tracer.cu, per-thread trace_closest / trace_any)
    against its CPU restatement over the oracle: same LCG streams, same operation order -> identical accumulation image, seeds and ray counts."""
    import luisa_compute_rs_b200.examples as ex
    desc = scenes.c2_cornell()
    w, h, spp, dispatches = 160, 128, 8, 3
    pt = ex.PathTracer(device, desc.meshes, w, h)
    o = ol.scene_from_desc(desc)
    img = np.zeros((h, w, 4), np.float32); seeds = ex.seed_image(w, h)
    cpu_rays = [0, 0]
    for _ in range(dispatches):
        pt.dispatch(spp, ex.MAX_DEPTH)
        c = ol.path_tracer_dispatch(o, desc.meshes, img, seeds, w, h, spp, ex.MAX_DEPTH, ex.TAN_HALF_FOV)
        cpu_rays[0] += c[0]; cpu_rays[1] += c[1]
    got_img, got_seeds = pt.download()
    assert np.array_equal(got_seeds, seeds)
    assert pt.rays == cpu_rays and cpu_rays[0] > w * h * spp * dispatches
    assert np.array_equal(got_img.view(np.uint32), img.view(np.uint32)), f"{(got_img != img).any(axis=2).sum()} of {w * h} pixels differ"
    rgb = got_img[..., :3] / got_img[..., 3:4]
    assert np.all(got_img[..., 3] == dispatches) and 0.05 < rgb.mean() < 1.0 and not np.isnan(rgb).any()
    # depth knob (BASELINE.json quotes depth 5; the example's loop bound is 10): still identical
    pt2 = ex.PathTracer(device, desc.meshes, 64, 64)
    img2 = np.zeros((64, 64, 4), np.float32); seeds2 = ex.seed_image(64, 64)
    pt2.dispatch(4, 5); ol.path_tracer_dispatch(o, desc.meshes, img2, seeds2, 64, 64, 4, 5, ex.TAN_HALF_FOV)
    assert np.array_equal(pt2.download()[0].view(np.uint32), img2.view(np.uint32))
    pt.destroy(); pt2.destroy(); o.close()


def test_stream_ordering_events_and_callbacks(device):
    s1, s2 = device.create_stream(), device.create_stream()
    a = device.create_buffer(1 << 16, 4); b = device.create_buffer(1 << 16, 4)
    src = np.arange(1 << 16, dtype=np.uint32)
    dst = np.zeros_like(src)
    fired = []
    ev = device.create_event()
    s1.submit([a.view().copy_from_async(src), a.view().copy_to_buffer_async(b.view())], callback=lambda: fired.append(1))
    ev.signal(s1, 1)
    ev.wait(s2, 1)
    s2.submit([b.view().copy_to_async(dst)], callback=lambda: fired.append(2))
    s2.synchronize(); s1.synchronize()
    assert np.array_equal(src, dst) and sorted(fired) == [1, 2] and ev.is_completed(1)
    ev.synchronize(1)
    assert np.array_equal(device.create_buffer(100, 12).view().to_numpy(np.uint8), np.zeros(1200, np.uint8))  # zero-initialised
    s1.destroy(); s2.destroy(); ev.destroy(); a.destroy(); b.destroy()


def test_timeline_events_wait_before_signal_and_monotone_counter(device):
    """EventImpl semantics (cpu/resource.rs:10-44, cpu/mod.rs:367-402): wait is enqueued, never blocks the caller, so a wait issued
    before its signal — from the same host thread — completes once the signal arrives; the counter is a fetch_max; value 0 of a fresh
    event is complete; thousands of signals on one event leave nothing behind."""
    s1, s2 = device.create_stream(), device.create_stream()
    ev = device.create_event()
    assert ev.is_completed(0) and not ev.is_completed(1)
    ev.synchronize(0)                                   # returns at once on a fresh event
    a = device.create_buffer(1 << 20, 4); b = device.create_buffer(1 << 20, 4)
    src = np.arange(1 << 20, dtype=np.uint32); dst = np.zeros_like(src)
    ev.wait(s2, 5)                                      # wait first ...
    s2.submit([b.view().copy_to_async(dst)])
    assert not ev.is_completed(5)
    s1.submit([a.view().copy_from_async(src), a.view().copy_to_buffer_async(b.view())])
    ev.signal(s1, 5)                                    # ... signal afterwards, same host thread
    s2.synchronize()
    assert np.array_equal(src, dst) and ev.is_completed(5) and ev.is_completed(3) and not ev.is_completed(6)
    ev.signal(s1, 2)                                    # a smaller value never lowers the counter
    s1.synchronize()
    assert ev.is_completed(5)
    for v in range(6, 3006):                            # per-frame signalling: one counter, no per-signal objects
        ev.signal(s1, v)
        if v % 500 == 0:
            ev.wait(s2, v)
    ev.synchronize(3005)
    assert ev.is_completed(3005) and not ev.is_completed(3006)
    s1.synchronize(); s2.synchronize()
    s1.destroy(); s2.destroy(); ev.destroy(); a.destroy(); b.destroy()


def test_upload_sources_are_snapshotted_before_dispatch_returns(device):
    """BufferUpload only borrows its source until dispatch() returns (cpu/stream.rs:33-64 copies it into staging buffers): the caller may
    overwrite or free it right after submit — also when it is pinned memory, which a plain cudaMemcpyAsync would still be reading."""
    import torch
    n = 1 << 24                                             # 64 MiB: long enough for a late DMA read to be caught
    stream = device.create_stream()
    buf = device.create_buffer(n, 4)
    for pinned in (True, False):
        host = torch.arange(n, dtype=torch.int32)
        if pinned:
            host = host.pin_memory()
        src = host.numpy().view(np.uint32)
        stream.submit([buf.view().copy_from_async(src)])
        src[:] = 0xDEADBEEF                                 # the borrow has ended
        stream.synchronize()
        got = buf.view().to_numpy(np.uint32)
        assert np.array_equal(got, np.arange(n, dtype=np.uint32)), f"pinned={pinned}: {(got != np.arange(n, dtype=np.uint32)).sum()} words changed under the copy"
    buf.destroy(); stream.destroy()


def test_builds_are_enqueued_and_do_not_block_the_submitting_thread(device):
    """MeshBuild / AccelBuild are stream work: dispatch() returns while the build is still running (the reference enqueues and returns,
    cpu/mod.rs:168-180; cuda_primitive.cpp:20-110 builds on the stream).  Measured with a host timestamp against the device time of
    the same build; the LBVH pipeline (AccelUsageHint::FastBuild) never reads anything back."""
    import time
    verts, tris = scenes.random_soup(4_000_000, 77)
    vb, ib = device.create_buffer_from_array(verts), device.create_buffer_from_array(tris)
    opt = lc.AccelOption(hint=lc.AccelUsageHint.FAST_BUILD)
    mesh = device.create_mesh(vb.view(), ib.view(), opt)
    accel = device.create_accel(opt)
    accel.push_mesh(mesh)
    s = device.create_stream()
    s.submit([mesh.build_async(), accel.build_async()]); s.synchronize()      # first build: allocations, arena growth
    first_nodes = mesh.stats()["wide_node_count"]
    times = []
    for _ in range(3):
        t0 = time.perf_counter()
        s.submit([mesh.build_async(), accel.build_async()])
        times.append(time.perf_counter() - t0)
        s.synchronize()
    st = mesh.stats()
    assert st["wide_node_count"] == first_nodes and st["primitive_count"] == 4_000_000 and st["was_refit"] == 0
    assert min(times) * 1e3 < 0.5 * st["build_ms"], f"submit took {min(times) * 1e3:.3f} ms, the build {st['build_ms']:.3f} ms on the device"
    # a refit is enqueued the same way
    mesh2 = device.create_mesh(vb.view(), ib.view(), lc.AccelOption(hint=lc.AccelUsageHint.FAST_BUILD, allow_update=True))
    s.submit([mesh2.build_async()]); s.synchronize(); mesh2.stats()
    s.submit([mesh2.build_async(lc.AccelBuildRequest.PREFER_UPDATE)]); s.synchronize(); mesh2.stats()   # first refit allocates its side arrays
    t0 = time.perf_counter(); s.submit([mesh2.build_async(lc.AccelBuildRequest.PREFER_UPDATE)]); dt = time.perf_counter() - t0
    s.synchronize()
    st2 = mesh2.stats()
    assert st2["was_refit"] == 1 and dt * 1e3 < 0.5 * st2["build_ms"], (dt * 1e3, st2["build_ms"])
    # results of an asynchronously built scene: the usual parity check on a sample
    rays = scenes.incoherent_rays(20000, seed=5)
    rb, hb = device.create_buffer_from_array(rays), device.create_buffer(20000, 24, 8)
    accel.intersect(rb, hb, 20000, 0xFF, s); s.synchronize()
    o = ol.OracleScene(); o.update(1, [{"index": 0, "flags": 1 | 2 | 4 | 16, "visibility": 0xFF, "mesh": o.add_mesh(verts, tris)}])
    assert_hits_equal(hb.view().to_numpy(lc.SurfaceHit), o.trace_closest(rays), "async build")
    o.close()
    for r in (rb, hb, accel, mesh, mesh2, vb, ib, s):
        r.destroy()


def test_update_instance_buffer_only_edits_the_table_and_keeps_the_tlas(device):
    """AccelBuildCommand.update_instance_buffer_only (api_types:643-652; AccelImpl::update returns before the scene commit,
    cpu/accel.rs:428-430; cuda_accel.cpp:277 skips the BVH build): the instance table takes the modifications — visibility, user id and
    opacity are read from it per instance entry, so they hold for the next traversal — while the TLAS keeps its boxes until the next
    full AccelBuild."""
    desc = scenes.instanced_scene(500, 4)
    d = DeviceScene(device, desc)
    o = ol.scene_from_desc(desc)
    rays = scenes.incoherent_rays(50000, lo=-1.0, hi=8.0, seed=13)
    tlas_before = d.accel.stats()
    assert_hits_equal(d.trace_closest(rays), o.trace_closest(rays), "before")
    d.accel.set_visibility_on_update(1, 0x0)
    d.accel.set_user_id_on_update(2, 4242)
    d.accel.build(instance_buffer_only=True)
    o.update(4, [{"index": 1, "flags": 16, "visibility": 0x0}, {"index": 2, "flags": 32, "user_id": 4242}])
    got = d.trace_closest(rays)
    assert_hits_equal(got, o.trace_closest(rays), "after update_instance_buffer_only")
    assert not (got["inst"] == 1).any() and d.accel.instance_user_id(2) == 4242 and d.accel.instance_visibility_mask(1) == 0
    assert d.accel.stats()["wide_node_count"] == tlas_before["wide_node_count"]      # the TLAS was not rebuilt
    # a transform edited the same way moves the instance's rays at once but its TLAS box only at the next full build
    t = np.eye(4, dtype=np.float32); t[:3, :] = desc.instances[3]["transform"]; t[1, 3] += 0.25
    d.accel.set_transform_on_update(3, t)
    d.accel.build(instance_buffer_only=True)
    d.accel.build()
    o.update(4, [{"index": 3, "flags": 2, "affine": t[:3, :].reshape(-1)}])
    assert_hits_equal(d.trace_closest(rays), o.trace_closest(rays), "after the full build")
    d.destroy(); o.close()


def test_counted_traversal_matches_and_reports_work(device):
    desc = scenes.c3_soup(20000, seed=61)
    rays = scenes.incoherent_rays(50000, seed=62)
    d = DeviceScene(device, desc)
    o = ol.scene_from_desc(desc)
    rb = device.create_buffer_from_array(rays); hb = device.create_buffer(rays.shape[0], 24, 8)
    c = d.accel.intersect_counted(rb, hb)
    assert_hits_equal(hb.view().to_numpy(lc.SurfaceHit), o.trace_closest(rays), "counted")
    assert c["rays"] == rays.shape[0] and c["nodes_visited"] > rays.shape[0] and c["tris_tested"] > 0
    st = d.meshes[0].stats()
    assert st["primitive_count"] == 20000 and st["packed_tri_count"] == 20000 and 0 < st["wide_node_count"] < 20000
    assert lc._abi.load_library().lc_b200_kernel_launch_count() > 0
    rb.destroy(); hb.destroy(); d.destroy(); o.close()


def test_full_size_properties_c3(device):
    """At BASELINE size (1M triangles, 16M rays) the oracle is too slow; check size-independent properties instead:
    any-hit == (closest-hit found), reversed-ray reciprocity of t, idempotence, and a 64k-ray sample against the oracle."""
    n_tris, n_rays = 1_000_000, 1 << 24
    desc = scenes.c3_soup(n_tris)
    d = DeviceScene(device, desc)
    rays = scenes.incoherent_rays(n_rays)
    h1 = d.accel.intersect_host(rays)
    h2 = d.accel.intersect_host(rays)
    assert h1.tobytes() == h2.tobytes()                                   # deterministic / idempotent
    occ = d.accel.intersect_any_host(rays)
    hit = h1["inst"] != lc.INVALID
    assert np.array_equal(occ != 0, hit)                                  # any-hit consistent with closest-hit
    assert 0.5 < hit.mean() < 1.0
    assert np.all(h1["committed_ray_t"][hit] > rays["tmin"][hit]) and np.all(h1["prim"][hit] < n_tris) and np.all(h1["inst"][hit] == 0)
    b = h1["bary"][hit]
    assert np.all(b >= -1e-6) and np.all(b.sum(1) <= 1 + 1e-6)
    # clipping the ray just short of the hit must miss that primitive; clipping at the hit must keep it
    sub = np.nonzero(hit)[0][:200000]
    r2 = rays[sub].copy(); r2["tmax"] = h1["committed_ray_t"][sub]
    assert np.array_equal(d.accel.intersect_host(r2)["prim"], h1["prim"][sub])
    r2["tmax"] = np.nextafter(h1["committed_ray_t"][sub], np.float32(0))
    assert (d.accel.intersect_host(r2)["inst"] == lc.INVALID).all()
    o = ol.scene_from_desc(desc)
    pick = np.random.default_rng(1).integers(0, n_rays, 65536)
    assert_hits_equal(h1[pick], o.trace_closest(rays[pick]), "1M-triangle sample")
    d.destroy(); o.close()


@pytest.mark.parametrize("builder", [0, 1, 2], ids=["lbvh", "ploc", "auto"])
def test_hits_do_not_depend_on_the_builder(device, builder):
    """LBVH split rule, PLOC clustering or the per-mesh choice: the canonical arithmetic makes hits tree-independent, so every
    builder must reproduce the oracle bit for bit — soup, terrain (PLOC's home turf), thousands of identical boxes (PLOC's
    degenerate case: equal distances everywhere), tiny meshes and an instanced scene."""
    lib = lc._abi.load_library()
    prev = lib.lc_b200_set_builder(builder)
    try:
        check_scene(device, scenes.c3_soup(30000), scenes.incoherent_rays(60000, seed=91))
        s = scenes.SceneDesc(); s.add_instance(s.add_mesh(*scenes.terrain(160)))
        rays = scenes.incoherent_rays(60000, seed=92); rays["orig"][:, 1] = rays["orig"][:, 1] * 0.3 + 0.05
        check_scene(device, s, rays)
        rng = np.random.default_rng(93)
        n = 5000
        v = np.tile(np.float32([[0, 0, 0], [1, 0, 0], [0, 1, 0]]), (n, 1, 1)).astype(np.float32)
        v[:, :, 2] = (rng.integers(0, 3, n) * np.float32(0.5))[:, None]
        s = scenes.SceneDesc(); s.add_instance(s.add_mesh(v.reshape(-1, 3), np.arange(3 * n, dtype=np.uint32).reshape(-1, 3)))
        o = np.concatenate([rng.random((5000, 2), dtype=np.float32) * 0.5, np.full((5000, 1), -1, np.float32)], 1)
        check_scene(device, s, scenes.make_rays(o, np.tile(np.float32([0, 0, 1]), (5000, 1)), 0.0, 10.0))
        for k in (1, 2, 3, 5, 17):
            check_scene(device, scenes.c3_soup(k, seed=94 + k), scenes.incoherent_rays(3000, seed=95))
        check_scene(device, scenes.instanced_scene(1500, 10, seed=96), scenes.incoherent_rays(40000, seed=97, lo=-1.0, hi=7.0))
        # points on a line with geometrically growing gaps: one mutual pair per PLOC iteration until the forced pairing takes over
        m = 3000
        x = np.cumsum(np.float32(1.01) ** np.arange(m, dtype=np.float32)).astype(np.float32)
        x /= x[-1]
        v = np.zeros((m, 3, 3), np.float32)
        v[:, :, 0] = x[:, None]; v[:, 1, 1] = 1e-3; v[:, 2, 2] = 1e-3
        s = scenes.SceneDesc(); s.add_instance(s.add_mesh(v.reshape(-1, 3), np.arange(3 * m, dtype=np.uint32).reshape(-1, 3)))
        r2 = scenes.incoherent_rays(20000, seed=98); r2["orig"][:, 1:] *= np.float32(2e-3); r2["dir"][:, 0] *= np.float32(0.05)
        check_scene(device, s, r2)
    finally:
        lib.lc_b200_set_builder(prev)


def _camera_rays(w, h, origin, look_at, fov_deg=45.0):
    o = np.asarray(origin, np.float32)
    f = np.asarray(look_at, np.float32) - o; f /= np.linalg.norm(f)
    r = np.cross(f, np.float32([0, 1, 0])); r /= np.linalg.norm(r)
    u = np.cross(r, f)
    t = np.float32(np.tan(np.radians(fov_deg) / 2))
    x = ((np.arange(w, dtype=np.float32) + 0.5) / w * 2 - 1) * t * np.float32(w / h)
    y = (1 - (np.arange(h, dtype=np.float32) + 0.5) / h * 2) * t
    d = (f[None, None, :] + x[None, :, None] * r[None, None, :] + y[:, None, None] * u[None, None, :]).reshape(-1, 3).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    return scenes.make_rays(np.broadcast_to(o, d.shape), d, np.float32(1e-4), np.float32(1e30))


def test_full_size_properties_c4_terrain_refit_and_rebuild(device):
    """Config C4 at BASELINE size (3164 x 3164 vertices = 20,009,138 triangles, 3840 x 2160 primary rays): a PreferUpdate refit of the
    displaced terrain must give exactly the hits of a fresh ForceBuild on the same vertices (hits are tree-independent), under both
    builders; any-hit == closest-hit found; and a ray sample equals the oracle on the full mesh."""
    lib = lc._abi.load_library()
    verts, tris = scenes.terrain(3164)
    assert tris.shape[0] == 20_009_138
    rays = _camera_rays(3840, 2160, [0.5, 0.6, -0.6], [0.5, 0.0, 0.5])
    n = rays.shape[0]
    vb = device.create_buffer_from_array(verts); ib = device.create_buffer_from_array(tris)
    mesh = device.create_mesh(vb.view(), ib.view(), lc.AccelOption(allow_update=True))
    accel = device.create_accel(lc.AccelOption(allow_update=True)); accel.push_mesh(mesh)
    rb = device.create_buffer_from_array(rays); hb = device.create_buffer(n, 24, 8); ob = device.create_buffer(n, 4, 4)

    def trace():
        accel.intersect(rb, hb, n); device.default_stream().synchronize()
        return hb.view().to_numpy(lc.SurfaceHit)
    moved = verts.copy(); moved[:, 1] += np.float32(0.01) * np.sin(np.float32(40.0) * verts[:, 0] + np.float32(0.3)).astype(np.float32)
    results = {}
    for builder in (0, 1):
        prev = lib.lc_b200_set_builder(builder)
        try:
            vb.view().copy_from(verts)
            mesh.build(lc.AccelBuildRequest.FORCE_BUILD); accel.build()
            results[builder, "base"] = trace()
            vb.view().copy_from(moved)
            mesh.build(lc.AccelBuildRequest.PREFER_UPDATE); assert mesh.stats()["was_refit"] == 1
            accel.build(lc.AccelBuildRequest.PREFER_UPDATE)
            refit = trace()
            mesh.build(lc.AccelBuildRequest.FORCE_BUILD); assert mesh.stats()["was_refit"] == 0
            accel.build()
            results[builder, "moved"] = trace()
            assert refit.tobytes() == results[builder, "moved"].tobytes(), "refit and rebuild disagree"
        finally:
            lib.lc_b200_set_builder(prev)
    assert results[0, "base"].tobytes() == results[1, "base"].tobytes() and results[0, "moved"].tobytes() == results[1, "moved"].tobytes()
    h = results[1, "moved"]
    hit = h["inst"] != lc.INVALID
    assert 0.2 < hit.mean() < 0.6 and (results[0, "base"]["prim"] != h["prim"]).any()
    accel.intersect_any(rb, ob, n); device.default_stream().synchronize()
    assert np.array_equal(ob.view().to_numpy(np.uint32) != 0, hit)
    desc = scenes.SceneDesc(); desc.add_instance(desc.add_mesh(moved, tris))
    o = ol.scene_from_desc(desc)
    pick = np.random.default_rng(5).integers(0, n, 30000)
    assert_hits_equal(h[pick], o.trace_closest(rays[pick]), "20M-triangle terrain sample")
    o.close()
    for r in (rb, hb, ob):
        r.destroy()
    accel.destroy(); mesh.destroy(); vb.destroy(); ib.destroy()


def test_full_size_properties_c5_instanced(device):
    """Config C5 at BASELINE size: one 4,999,122-triangle terrain instanced 10x (yaw 36 deg * k on a 5 x 2 grid), 4K primary rays:
    determinism, any-hit consistency, every hit inside its instance's triangle range, and a ray sample against the oracle."""
    verts, tris = scenes.terrain(1582)
    desc = scenes.SceneDesc()
    mid = desc.add_mesh(verts, tris)
    for k in range(10):
        t = scenes.rotation_y(36.0 * k); t[:, 3] = [1.2 * (k % 5), 0.0, 1.2 * (k // 5)]
        desc.add_instance(mid, t)
    assert desc.triangle_count() == 49_991_220
    d = DeviceScene(device, desc)
    rays = _camera_rays(3840, 2160, [3.0, 2.5, -3.0], [3.0, 0.0, 1.0])
    h1 = d.trace_closest(rays); h2 = d.trace_closest(rays)
    assert h1.tobytes() == h2.tobytes()
    hit = h1["inst"] != lc.INVALID
    assert np.array_equal(d.trace_any(rays) != 0, hit) and 0.15 < hit.mean() < 0.6
    assert np.all(h1["inst"][hit] < 10) and np.all(h1["prim"][hit] < tris.shape[0]) and len(np.unique(h1["inst"][hit])) == 10
    o = ol.scene_from_desc(desc)
    pick = np.random.default_rng(6).integers(0, rays.shape[0], 30000)
    assert_hits_equal(h1[pick], o.trace_closest(rays[pick]), "50M-triangle instanced sample")
    d.destroy(); o.close()


@pytest.mark.gpu
def test_upload_right_after_create_buffer_survives_the_zero_fill(device):
    """create_buffer zero-fills (BufferImpl::new, cpu/resource.rs:126-133) on the legacy stream; the device's streams are
    non-blocking, so the fill of a large buffer must be over before create_buffer returns or it overtakes the first upload."""
    rng = np.random.default_rng(5)
    for _ in range(4):
        big = device.create_buffer(64 << 20, 4, 4)  # 256 MiB: a fill that takes a while
        head = rng.integers(0, 2**32, 4096, dtype=np.uint32)
        big.view(0, head.shape[0]).copy_from(head)
        small = device.create_buffer_from_array(head)
        back = np.zeros_like(head); big.view(0, head.shape[0]).copy_to(back)
        back2 = np.zeros_like(head); small.view().copy_to(back2)
        tail = np.ones(16, np.uint32); big.view((64 << 20) - 16, 16).copy_to(tail)
        assert np.array_equal(back, head) and np.array_equal(back2, head) and not tail.any()
        big.destroy(); small.destroy()
