"""Development bench for the DSL call path (create_shader + ShaderDispatch) against the batch entry point, with the knobs of the
wavefront lowering swept in-process.  usage: python tools/dsl_bench.py [c3] [c2] [--rays N]"""
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import luisa_compute_rs_b200 as lc  # noqa: E402
import scenes  # noqa: E402
from luisa_compute_rs_b200 import examples_ir  # noqa: E402

lib = lc._abi.load_library()


def timed(stream, ext, fn, reps):
    fn(); stream.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(ext)
    for _ in range(reps):
        fn()
    e1.record(ext); stream.synchronize()
    return e0.elapsed_time(e1) / reps


def c3(dev, n_rays, configs):
    verts, tris = scenes.random_soup(1_000_000, 0x5EED0001)
    vb, ib = dev.create_buffer_from_array(verts), dev.create_buffer_from_array(tris)
    mesh = dev.create_mesh(vb.view(), ib.view(), lc.AccelOption()); mesh.build(lc.AccelBuildRequest.FORCE_BUILD)
    accel = dev.create_accel(); accel.push_mesh(mesh); accel.build()
    rays = scenes.incoherent_rays(n_rays, seed=0x5EED0002)
    rb = dev.create_buffer(n_rays, 32, 16); rb.view().copy_from(rays)
    hb, hb2 = dev.create_buffer(n_rays, 24, 8), dev.create_buffer(n_rays, 24, 8)
    stream = dev.create_stream()
    ext = torch.cuda.ExternalStream(stream.cuda_stream())
    ms = timed(stream, ext, lambda: accel.intersect(rb, hb2, n_rays, 0xFF, stream), 5)
    print(json.dumps({"c3": "batch k_trace", "ms": ms, "mrays": n_rays / ms / 1e3}), flush=True)
    ref = hb2.view().to_numpy(lc.SurfaceHit).tobytes()
    for cfg in configs:
        mode, minb, yld = cfg[:3]
        lib.lc_b200_set_lowering(mode)
        os.environ["LC_B200_WAVE_MIN_BLOCKS"] = str(minb); os.environ["LC_B200_WAVE_YIELD"] = str(yld)
        k = examples_ir.trace_buffer_kernel()
        t0 = time.perf_counter()
        sh = dev.create_shader(C.addressof(k.km), keep=k)
        cs = time.perf_counter() - t0
        ms = timed(stream, ext, lambda: stream.submit([sh.dispatch_async((n_rays, 1, 1), rb, hb, accel)]), 5)
        same = hb.view().to_numpy(lc.SurfaceHit).tobytes() == ref
        print(json.dumps({"c3": "dsl", "lowering": ["wavefront", "direct"][mode == 1], "min_blocks": minb, "yield": yld, "ms": ms, "mrays": n_rays / ms / 1e3,
                          "identical": same, "create_s": cs}), flush=True)
        sh.destroy()
    lib.lc_b200_set_lowering(0)


def c2(dev, configs):
    import luisa_compute_rs_b200.examples as ex
    w = h = 1024
    desc = scenes.c2_cornell()
    pt = ex.PathTracer(dev, desc.meshes, w, h)
    for _ in range(2):
        pt.dispatch(32, 10, count_rays=True)
    rays_per_dispatch = sum(pt.rays) / 2
    n = len(desc.meshes)
    vheap, iheap = dev.create_bindless_array(n), dev.create_bindless_array(n)
    for i, (vb, ib) in enumerate(zip(pt.vbuffers, pt.ibuffers)):
        vheap.emplace_buffer_async(i, vb); iheap.emplace_buffer_async(i, ib)
    s = dev.default_stream()
    s.submit([vheap.update_async(), iheap.update_async()])
    sext = torch.cuda.ExternalStream(s.cuda_stream())
    res = np.array([w, h], np.uint32)
    ref = None
    for cfg in configs:
        mode, minb, yld = cfg[:3]
        fast = bool(cfg[3]) if len(cfg) > 3 else False
        lib.lc_b200_set_lowering(mode)
        os.environ["LC_B200_WAVE_MIN_BLOCKS"] = str(minb); os.environ["LC_B200_WAVE_YIELD"] = str(yld)
        image = dev.create_tex2d("Rgba32f", w, h); seeds = dev.create_tex2d("R32Uint", w, h)
        seeds.copy_from(ex.seed_image(w, h).reshape(h, w))
        k = examples_ir.path_tracer_kernel(vheap.handle.id, iheap.handle.id, 32, 10, polynomial_sincos=not fast)
        sh = dev.create_shader(C.addressof(k.km), fast_math=fast, keep=k)
        sh.dispatch((w, h), image, seeds, pt.accel, res)
        img = image.to_numpy().tobytes()
        if ref is None:
            ref = img
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(sext)
        s.submit([sh.dispatch_async((w, h), image, seeds, pt.accel, res) for _ in range(3)])
        e1.record(sext); s.synchronize()
        ms = e0.elapsed_time(e1) / 3
        print(json.dumps({"c2": "dsl path tracer", "lowering": ["wavefront", "direct"][mode == 1], "fast_math": fast, "min_blocks": minb, "yield": yld, "ms_per_dispatch": ms,
                          "mrays": rays_per_dispatch / ms / 1e3, "first_dispatch_identical": img == ref}), flush=True)
        for r in (sh, image, seeds):
            r.destroy()
    lib.lc_b200_set_lowering(0)


def main():
    args = sys.argv[1:]
    n_rays = int(args[args.index("--rays") + 1]) if "--rays" in args else 1 << 24
    ctx = lc.Context(); dev = ctx.create_device("b200")
    wave = [(0, mb, y) for mb in (4, 5, 6, 7) for y in (4, 8, 16)]
    only = None
    if "--configs" in args:   # "mode:min_blocks:yield,..."  (mode 0 wavefront, 1 direct)
        only = [tuple(int(v) for v in c.split(":")) for c in args[args.index("--configs") + 1].split(",")]
    if "c3" in args:
        c3(dev, n_rays, only or [(1, 4, 8)] + wave)
    if "c2" in args:
        c2(dev, only or [(1, 4, 8)] + [(0, mb, y) for mb in (3, 4, 5) for y in (4, 8, 16, 24, 32)])
    dev.close()


if __name__ == "__main__":
    main()
