"""The reference call path for ray queries — `accel.intersect(ray, mask)` inside a DSL kernel (rtx.rs:774-817), reaching the device through
create_shader + ShaderDispatch (cpu/stream.rs:330-409, codegen cpp.rs:1334-1472) — on both lowerings of csrc/ir_lower.cpp:

* wavefront (default for kernels that trace from their own body): persistent threads, trace calls are suspension points of one
  warp-synchronous traversal loop shared with the batch kernel;
* direct: one dispatch id per CUDA thread, per-thread traversal.

Both must return exactly what the oracle returns (and therefore what the batch entry points return): the arithmetic per ray is fixed,
only the schedule differs.  CPU part: the host-side launch geometry of the wavefront form (work item -> thread / block / dispatch id).
"""
import ctypes as C

import numpy as np
import pytest

import luisa_compute_rs_b200 as lc
import oracle_lib as ol
import scenes
from harness import DeviceScene, assert_hits_equal
from luisa_compute_rs_b200 import examples_ir, ir

AUTO, DIRECT = 0, 1


class lowering:
    def __init__(self, mode): self.mode = mode
    def __enter__(self): self.prev = lc._abi.load_library().lc_b200_set_lowering(self.mode)
    def __exit__(self, *a): lc._abi.load_library().lc_b200_set_lowering(self.prev)


def wave_ids(item, dispatch_size, block_size):
    """numpy restatement of lc_wave_ids (lc_device_lib.cuh): the order in which a wavefront-lowered kernel hands out dispatch ids"""
    bx, by, bz = block_size
    gx, gy = -(-dispatch_size[0] // bx), -(-dispatch_size[1] // by)
    b, t = divmod(item, bx * by * bz)
    bzi, bxy = divmod(b, gx * gy)
    block = (bxy % gx, bxy // gx, bzi)
    thread = (t % bx, (t // bx) % by, t // (bx * by))
    return tuple(block[k] * block_size[k] + thread[k] for k in range(3))


def test_work_items_enumerate_every_dispatch_id_once():
    for ds, bs in (((70, 33, 3), (16, 8, 2)), ((1, 1, 1), (64, 1, 1)), ((129, 1, 1), (64, 1, 1)), ((5, 7, 1), (16, 16, 1))):
        grid = [-(-ds[k] // bs[k]) for k in range(3)]
        n_items = grid[0] * grid[1] * grid[2] * bs[0] * bs[1] * bs[2]
        ids = [wave_ids(i, ds, bs) for i in range(n_items)]
        live = [d for d in ids if all(d[k] < ds[k] for k in range(3))]
        assert len(live) == ds[0] * ds[1] * ds[2] and len(set(live)) == len(live)
        # consecutive items of one block stay inside that block's tile (what keeps a warp's rays neighbours)
        assert all(ids[i][0] // bs[0] == ids[0][0] // bs[0] for i in range(bs[0] * bs[1] * bs[2]))


def k_ids_kernel(block_size):
    """writes thread_id / block_id / dispatch_id per dispatch id and traces one ray in between, so that the ids live across a yield"""
    k = ir.KernelBuilder(block_size=block_size)
    _, ray_ty, hit_ty = examples_ir.common_types(k)
    out = k.arg_buffer(k.u32)
    accel = k.arg_accel()
    size = k.arg_uniform(k.u323)
    f3 = k.array(k.f32, 3)

    def body():
        d, t, b = k.dispatch_id(), k.thread_id(), k.block_id()
        lin = (d.z * size.y + d.y) * size.x + d.x
        o = k.vec(k.f323, d.x.cast(k.f32) * 0.01, d.y.cast(k.f32) * 0.01, -1.0)
        ray = examples_ir.make_ray(k, ray_ty, f3, o, k.vec(k.f323, 0.0, 0.0, 1.0), 0.0, 100.0)
        hit = accel.trace_closest(ray, 0xFF, hit_ty)
        base = lin * k.u(10)
        for j, v in enumerate((d.x, d.y, d.z, t.x, t.y, t.z, b.x, b.y, b.z)):
            out.write(base + k.u(j), v)
        out.write(base + k.u(9), hit.extract(1))
    k.body(body)
    k.finish()
    return k


def test_ids_kernel_compiles_in_both_lowerings():
    lib = lc._abi.load_library()
    for mode in (AUTO, DIRECT):
        with lowering(mode):
            k = k_ids_kernel((8, 4, 2))
            log = C.c_void_p()
            assert lib.lc_b200_shader_compile_check(C.addressof(k.km), False, C.byref(log)) == 0, C.string_at(log).decode()


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [AUTO, DIRECT], ids=["wavefront", "direct"])
def test_thread_block_and_dispatch_ids_across_a_trace_call(device, mode):
    desc = scenes.c1_triangle()
    d = DeviceScene(device, desc)
    ds, bs = (37, 13, 3), (8, 4, 2)
    n = ds[0] * ds[1] * ds[2]
    out = device.create_buffer(n * 10, 4, 4)
    with lowering(mode):
        k = k_ids_kernel(bs)
        sh = device.create_shader(C.addressof(k.km), keep=k)
    sh.dispatch(ds, out, d.accel, np.array(list(ds) + [0], np.uint32))
    got = out.view().to_numpy(np.uint32).reshape(ds[2], ds[1], ds[0], 10)
    z, y, x = np.meshgrid(np.arange(ds[2]), np.arange(ds[1]), np.arange(ds[0]), indexing="ij")
    want = np.stack([x, y, z, x % bs[0], y % bs[1], z % bs[2], x // bs[0], y // bs[1], z // bs[2]], -1).astype(np.uint32)
    assert np.array_equal(got[..., :9], want)
    rays = np.zeros(n, lc.Ray)
    rays["orig"] = np.stack([x.reshape(-1) * np.float32(0.01), y.reshape(-1) * np.float32(0.01), np.full(n, -1.0)], -1).astype(np.float32)
    rays["dir"] = (0, 0, 1); rays["tmax"] = 100.0
    o = ol.scene_from_desc(desc)
    assert np.array_equal(got[..., 9].reshape(-1), o.trace_closest(rays)["prim"])
    sh.destroy(); out.destroy(); d.destroy(); o.close()


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [AUTO, DIRECT], ids=["wavefront", "direct"])
def test_dsl_trace_kernel_equals_the_oracle_and_the_batch_entry_points(device, mode):
    cases = [
        (scenes.c3_soup(20000, seed=3), scenes.incoherent_rays(100003, seed=4), (0xFF,)),
        (scenes.c2_cornell(), scenes.c2_primary_rays(96, 80), (0xFF,)),
        (scenes.instanced_scene(2000, 10), scenes.incoherent_rays(60001, lo=-1.0, hi=8.0, seed=31), (0xFF, 0xF0, 0x0)),
        (scenes.c3_soup(1, seed=5), scenes.incoherent_rays(129, seed=6), (0xFF,)),
    ]
    with lowering(mode):
        for desc, rays, masks in cases:
            o = ol.scene_from_desc(desc)
            d = DeviceScene(device, desc)
            for mask in masks:
                want = o.trace_closest(rays, mask, ol.BVH)
                got = d.trace_dsl(rays, mask=mask)
                assert_hits_equal(got, want, f"DSL kernel, mask {mask:#x}")
                assert got.tobytes() == d.trace_closest(rays, mask).tobytes()
                assert np.array_equal(d.trace_dsl(rays, any_hit=True, mask=mask), o.trace_any(rays, mask, ol.BVH))
            d.destroy(); o.close()


@pytest.mark.gpu
def test_dsl_trace_kernel_on_an_empty_accel_and_an_empty_dispatch(device):
    s = scenes.SceneDesc()
    d = DeviceScene(device, s)
    rays = scenes.incoherent_rays(1000, seed=8)
    got = d.trace_dsl(rays)
    assert (got["inst"] == lc.INVALID).all() and (got["prim"] == lc.INVALID).all() and np.array_equal(got["committed_ray_t"], rays["tmax"])
    assert (d.trace_dsl(rays, any_hit=True) == 0).all()
    assert d.trace_dsl(rays[:0]).shape == (0,)
    d.destroy()


@pytest.mark.gpu
def test_dsl_trace_kernel_full_size_c3_equals_the_batch_kernel(device):
    """Config C3 at BASELINE size through the reference call path: 16 Mi rays by create_shader + ShaderDispatch, byte-identical to the
    batch entry point (itself checked against the oracle in test_parity_gpu.py)."""
    n_rays = 1 << 24
    desc = scenes.c3_soup(1_000_000)
    d = DeviceScene(device, desc)
    rays = scenes.incoherent_rays(n_rays)
    rb = device.create_buffer(n_rays, 32, 16); rb.view().copy_from(rays)
    hb, hb2 = device.create_buffer(n_rays, 24, 8), device.create_buffer(n_rays, 24, 8)
    k = examples_ir.trace_buffer_kernel()
    sh = device.create_shader(C.addressof(k.km), keep=k)
    sh.dispatch((n_rays, 1, 1), rb, hb, d.accel)
    d.accel.intersect(rb.view(), hb2.view(), n_rays, 0xFF)
    device.default_stream().synchronize()
    a, b = hb.view().to_numpy(lc.SurfaceHit), hb2.view().to_numpy(lc.SurfaceHit)
    assert a.tobytes() == b.tobytes()
    assert 0.5 < (a["inst"] != lc.INVALID).mean() < 1.0
    for r in (sh, rb, hb, hb2):
        r.destroy()
    d.destroy()
