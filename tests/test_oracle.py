"""CPU tests of the oracle (the checker itself): closed-form golden vectors, brute force == BVH, reference semantics of
AccelImpl::update, byte layouts.  No GPU, no CUDA library calls."""
import os

import numpy as np
import pytest

import oracle_lib as ol
import scenes

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def test_layouts_match_reference_rtx_layout_test():
    # the reference's only hot-path test: rtx.rs:550-561 (Ray 32/16, SurfaceHit 24/8, Index 12)
    assert scenes.RAY.itemsize == 32 and scenes.HIT.itemsize == 24 and ol.HIT.itemsize == 24
    assert np.dtype(("<u4", (3,))).itemsize == 12


def test_c1_closed_form_golden():
    """C1 (raytracing.rs triangle): t = 1/d.z and barycentrics from the closed-form plane hit, computed in float64 by
    tests/golden/make_golden.py; the oracle's fp32 canonical arithmetic must agree within 1e-5 and on hit/miss away from edges."""
    g = np.load(os.path.join(GOLDEN, "c1_closed_form_64.npz"))
    rays = scenes.c1_rays(64, 64)
    o = ol.scene_from_desc(scenes.c1_triangle())
    hits = o.trace_closest(rays, mode=ol.BRUTE)
    clear = g["edge_margin"] > 1e-5
    hit = hits["inst"] != 0xFFFFFFFF
    assert np.array_equal(hit[clear], g["hit"][clear])
    v = hit & g["hit"] & clear
    assert v.sum() > 100
    assert np.all(hits["inst"][v] == 0) and np.all(hits["prim"][v] == 0)
    assert np.max(np.abs(hits["committed_ray_t"][v] - g["t"][v]) / g["t"][v]) < 1e-5
    assert np.max(np.abs(hits["bary"][v] - g["bary"][v])) < 1e-5
    miss = ~hit
    assert np.all(hits["committed_ray_t"][miss] == rays["tmax"][miss]) and np.all(hits["bary"][miss] == 0)


def test_golden_regression_vectors():
    """Committed oracle outputs (self-generated: the reference holds no golden hits, parity unpinned)."""
    g = np.load(os.path.join(GOLDEN, "cornell_primary_48.npz"))
    o = ol.scene_from_desc(scenes.c2_cornell())
    hits = o.trace_closest(scenes.c2_primary_rays(48, 48), mode=ol.BRUTE)
    assert np.array_equal(hits["inst"], g["inst"]) and np.array_equal(hits["prim"], g["prim"])
    assert np.array_equal(hits["committed_ray_t"].view(np.uint32), g["t_bits"])
    g = np.load(os.path.join(GOLDEN, "soup2k_rays4k.npz"))
    o = ol.scene_from_desc(scenes.c3_soup(2000))
    hits = o.trace_closest(scenes.incoherent_rays(4096), mode=ol.BRUTE)
    assert np.array_equal(hits["inst"], g["inst"]) and np.array_equal(hits["prim"], g["prim"])
    assert np.array_equal(hits["committed_ray_t"].view(np.uint32), g["t_bits"])


@pytest.mark.parametrize("n_tris,n_rays", [(1, 500), (7, 2000), (3000, 6000)])
def test_bvh_equals_brute_force(n_tris, n_rays):
    o = ol.scene_from_desc(scenes.c3_soup(n_tris, seed=n_tris))
    rays = scenes.incoherent_rays(n_rays, seed=n_rays, tmin=0.0)
    a, b, c = o.trace_closest(rays, mode=ol.BRUTE), o.trace_closest(rays, mode=ol.BVH), o.trace_closest(rays, mode=ol.WIDE)
    assert a.tobytes() == b.tobytes() == c.tobytes()      # brute force = binary BVH = 8-wide AVX2 traversal (the CPU baseline's fast path)
    assert np.array_equal(o.trace_any(rays, mode=ol.BRUTE), o.trace_any(rays, mode=ol.BVH))
    assert np.array_equal(o.trace_any(rays, mode=ol.BRUTE), o.trace_any(rays, mode=ol.WIDE))
    assert np.array_equal(o.trace_any(rays, mode=ol.BRUTE) != 0, a["inst"] != 0xFFFFFFFF)


def test_instanced_bvh_equals_brute_and_masks():
    desc = scenes.instanced_scene(600, 10)
    o = ol.scene_from_desc(desc)
    rays = scenes.incoherent_rays(5000, lo=-1.0, hi=8.0, seed=11)
    for mask in (0xFF, 0xF0, 0x01, 0):
        a, b = o.trace_closest(rays, mask, ol.BRUTE), o.trace_closest(rays, mask, ol.BVH)
        assert a.tobytes() == b.tobytes() == o.trace_closest(rays, mask, ol.WIDE).tobytes()
        hit = a["inst"] != 0xFFFFFFFF
        vis = np.array([i["mask"] for i in desc.instances], np.uint32)
        assert np.all((vis[a["inst"][hit]] & mask) != 0)
    assert (o.trace_closest(rays, 0, ol.BVH)["inst"] == 0xFFFFFFFF).all()


def test_fp32_vs_f64_truth_within_tolerance():
    o = ol.scene_from_desc(scenes.c3_soup(5000))
    rays = scenes.incoherent_rays(20000)
    hits = o.trace_closest(rays)
    truth, amb = o.truth(rays)
    clear = amb == 0
    assert np.array_equal(hits["inst"][clear], truth["inst"][clear]) and np.array_equal(hits["prim"][clear], truth["prim"][clear])
    v = clear & (truth["inst"] != 0xFFFFFFFF)
    assert np.max(np.abs(hits["committed_ray_t"][v] - truth["committed_ray_t"][v]) / truth["committed_ray_t"][v]) < 1e-5
    assert amb.mean() < 0.01


def test_ties_resolve_to_lowest_ids():
    # two coincident triangles in one mesh and the same mesh instanced twice at the same place
    s = scenes.SceneDesc()
    v = [[0, 0, 1], [1, 0, 1], [0, 1, 1]]
    m = s.add_mesh(v + v, [[3, 4, 5], [0, 1, 2]])
    s.add_instance(m); s.add_instance(m)
    o = ol.scene_from_desc(s)
    rays = scenes.make_rays([[0.2, 0.2, 0]], [[0, 0, 1]], 0.0, 10.0)
    for mode in (ol.BRUTE, ol.BVH):
        h = o.trace_closest(rays, mode=mode)
        assert (h["inst"][0], h["prim"][0]) == (0, 0) and abs(h["committed_ray_t"][0] - 1.0) < 1e-6
    assert o.trace_closest(rays, mode=ol.BVH, mask=0xFF)["inst"][0] == 0


def test_hit_interval_is_open_closed():
    s = scenes.SceneDesc()
    s.add_instance(s.add_mesh([[0, 0, 1], [1, 0, 1], [0, 1, 1]], [[0, 1, 2]]))
    o = ol.scene_from_desc(s)
    mk = lambda tmin, tmax: scenes.make_rays([[0.2, 0.2, 0]], [[0, 0, 1]], tmin, tmax)
    t = o.trace_closest(mk(0.0, 10.0))["committed_ray_t"][0]  # the canonical fp32 t (within an ulp of 1)
    assert abs(t - 1.0) < 1e-6
    below = np.nextafter(t, np.float32(0))
    assert o.trace_closest(mk(0.0, t))["inst"][0] == 0            # t == tmax accepted
    assert o.trace_closest(mk(t, 2.0))["inst"][0] == 0xFFFFFFFF    # t == tmin rejected
    assert o.trace_closest(mk(below, 2.0))["inst"][0] == 0
    assert o.trace_closest(mk(0.0, below))["inst"][0] == 0xFFFFFFFF
    assert o.trace_any(mk(0.0, t))[0] == 1 and o.trace_any(mk(t, 2.0))[0] == 0
    # back faces are not culled
    back = scenes.make_rays([[0.2, 0.2, 2]], [[0, 0, -1]], 0.0, 10.0)
    assert o.trace_closest(back)["inst"][0] == 0


def test_accel_update_semantics():
    """AccelImpl::update (cpu/accel.rs:324-447): PRIMITIVE resets mask/opaque and takes the affine as given; later flags apply
    in order; shrinking pops; empty slots are skipped."""
    o = ol.OracleScene()
    m = o.add_mesh(np.array([[0, 0, 1], [1, 0, 1], [0, 1, 1]], np.float32), np.array([[0, 1, 2]], np.uint32))
    shift = [1, 0, 0, 5, 0, 1, 0, 0, 0, 0, 1, 0]
    o.update(3, [dict(index=2, flags=1, mesh=m, user_id=9, affine=shift)])  # set_handle quirk: PRIMITIVE without TRANSFORM still moves it
    assert o.instance_user_id(2) == 9 and o.instance_visibility(2) == 0xFF
    assert np.array_equal(o.instance_transform(2).reshape(12), np.array(shift, np.float32))
    r0 = scenes.make_rays([[0.2, 0.2, 0]], [[0, 0, 1]], 0.0, 10.0)
    r5 = scenes.make_rays([[5.2, 0.2, 0]], [[0, 0, 1]], 0.0, 10.0)
    assert o.trace_closest(r0)["inst"][0] == 0xFFFFFFFF and o.trace_closest(r5)["inst"][0] == 2
    o.update(3, [dict(index=2, flags=16, visibility=0x2)])
    assert o.trace_closest(r5, mask=0x1)["inst"][0] == 0xFFFFFFFF and o.trace_closest(r5, mask=0x2)["inst"][0] == 2
    o.update(3, [dict(index=0, flags=1 | 2, mesh=m)])
    assert o.trace_closest(r0, mask=0xFF)["inst"][0] == 0
    o.update(1, [])
    assert o.trace_closest(r5, mask=0xFF)["inst"][0] == 0xFFFFFFFF and o.trace_closest(r0)["inst"][0] == 0


def test_offset_ray_origin_matches_host_mirror():
    import luisa_compute_rs_b200 as lc
    rng = np.random.default_rng(3)
    p = ((rng.random((2000, 3)) - 0.5) * np.array([0.05, 4.0, 100.0])).astype(np.float32)
    n = rng.normal(size=(2000, 3)).astype(np.float32)
    n /= np.linalg.norm(n, axis=1, keepdims=True)
    assert np.array_equal(ol.offset_ray_origin(p, n).view(np.uint32), lc.offset_ray_origin(p, n).view(np.uint32))


def test_degenerate_inputs():
    s = scenes.SceneDesc()
    # zero-area triangle, a sliver, and a regular one
    s.add_instance(s.add_mesh([[0, 0, 1], [0, 0, 1], [0, 0, 1], [0, 0, 2], [1, 0, 2], [2, 0, 2], [0, 0, 3], [1, 0, 3], [0, 1, 3]], [[0, 1, 2], [3, 4, 5], [6, 7, 8]]))
    o = ol.scene_from_desc(s)
    rays = scenes.make_rays([[0.1, 0.1, 0], [0, 0, 0], [0.5, 0, 0]], [[0, 0, 1], [0, 0, 0], [0, 0, 1]], 0.0, 10.0)
    a, b = o.trace_closest(rays, mode=ol.BRUTE), o.trace_closest(rays, mode=ol.BVH)
    assert a.tobytes() == b.tobytes()
    assert a["prim"][0] == 2 and a["inst"][1] == 0xFFFFFFFF  # zero direction never hits
    empty = ol.OracleScene(); empty.update(0, [])
    assert empty.trace_closest(rays)["inst"].tolist() == [0xFFFFFFFF] * 3


def test_ray_query_semantics():
    """RayQuery (accel.rs:582-800): opaque instances commit without a callback, candidates of non-opaque instances go through
    the hook; commit-all == trace_closest; reject-all hides non-opaque instances; BVH == brute force; nothing committed -> Miss."""
    s = scenes.SceneDesc()
    m = s.add_mesh(*scenes.random_soup(400, 5, extent=0.2))
    s.add_instance(m, opaque=False)
    t = scenes.IDENTITY34.copy(); t[:, 3] = [0.3, 0.1, 0]
    s.add_instance(m, t, opaque=True)
    o = ol.scene_from_desc(s)
    rays = scenes.incoherent_rays(4000, seed=3)
    closest = o.trace_closest(rays)
    q = o.ray_query(rays, kind=0)
    assert np.array_equal(q["inst"], closest["inst"]) and np.array_equal(q["prim"], closest["prim"])
    hit = q["hit_type"] == 1
    assert np.array_equal(hit, closest["inst"] != 0xFFFFFFFF)
    assert np.array_equal(q["committed_ray_t"][hit], closest["committed_ray_t"][hit]) and np.all(q["committed_ray_t"][~hit] == 0)
    assert np.array_equal(q["bary"][hit], closest["bary"][hit])
    rej = o.ray_query(rays, kind=3)
    assert np.all(rej["inst"][rej["hit_type"] == 1] == 1)        # only the opaque instance is ever committed
    only_opaque = o.trace_closest(rays, mask=0xFF)                  # reference: same scene traced with instance 0 removed
    s2 = scenes.SceneDesc(); m2 = s2.add_mesh(*scenes.random_soup(400, 5, extent=0.2)); s2.add_instance(m2, mask=0); s2.add_instance(m2, t)
    o2 = ol.scene_from_desc(s2)
    c2 = o2.trace_closest(rays)
    assert np.array_equal(rej["prim"], c2["prim"]) and np.array_equal(rej["hit_type"] == 1, c2["inst"] != 0xFFFFFFFF)
    for kind, kw in ((1, dict(radius=0.8)), (2, dict(bits=np.random.default_rng(1).integers(0, 2**32, 26, dtype=np.uint64).astype(np.uint32), first_bit=[0, 400]))):
        a = o.ray_query(rays, kind=kind, **kw)
        b = o.ray_query(rays, kind=kind, mode=ol.BRUTE, **kw)
        assert a.tobytes() == b.tobytes()
        n_a = (a["hit_type"] == 1).sum()
        assert (rej["hit_type"] == 1).sum() <= n_a <= hit.sum()
        anyq = o.ray_query(rays, terminate_on_first=True, kind=kind, **kw)
        assert np.array_equal(anyq["hit_type"], a["hit_type"])   # which hit is order dependent, whether one exists is not
    o.close(); o2.close()


def test_wide_mode_and_parallel_build_on_a_larger_scene():
    """mode 2 (8-wide tree, AVX2 box tests; what `bench.py --impl reference` times) returns the canonical hits of the scalar modes on a
    scene large enough for the multi-threaded SAH build, on grazing / axis-aligned rays and with a zero direction component."""
    verts, tris = scenes.terrain(200)
    s = scenes.SceneDesc(); s.add_instance(s.add_mesh(verts, tris), np.eye(4, dtype=np.float32)[:3])
    o = ol.scene_from_desc(s)
    rays = scenes.incoherent_rays(60000, seed=77)
    rays["orig"][:, 1] = rays["orig"][:, 1] * 0.3 + 0.05
    rays["dir"][:20000, 1] = 0.0                     # exactly horizontal: a zero direction component in the slab test
    rays["dir"][20000:30000] = (0.0, -1.0, 0.0)      # straight down
    a, b = o.trace_closest(rays, mode=ol.BVH), o.trace_closest(rays, mode=ol.WIDE)
    assert a.tobytes() == b.tobytes() and (a["inst"] != 0xFFFFFFFF).mean() > 0.3
    assert np.array_equal(o.trace_any(rays, mode=ol.BVH), o.trace_any(rays, mode=ol.WIDE))
    o.close()
