#!/usr/bin/env python
"""Config C2 (examples/path_tracer.rs: Cornell box 1024x1024, 256 spp = 8 dispatches x 32 spp) on one GPU, two ways:
  hand    : csrc/path_tracer.cu behind lc_b200_example_path_tracer
  lowered : the example's kernel as an ir::KernelModule -> create_shader (IR -> CUDA lowering + NVRTC) -> ShaderDispatch
Prints one JSON line per variant: ms per dispatch, Mrays/s (closest + any, counted by the hand kernel on the same seeds),
create_shader time, and whether the two accumulation images are bit-identical (polynomial sin/cos variant)."""
import argparse
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=1024)
    ap.add_argument("--spp", type=int, default=32)
    ap.add_argument("--dispatches", type=int, default=8)
    ap.add_argument("--depth", type=int, default=10)
    ap.add_argument("--cutout", action="store_true", help="also time examples/path_tracer_cutout.rs (ray queries with a candidate filter)")
    a = ap.parse_args()
    import torch
    import luisa_compute_rs_b200 as lc
    import luisa_compute_rs_b200.examples as ex
    from luisa_compute_rs_b200 import examples_ir
    import scenes
    dev = lc.Context().create_device("b200")
    desc = scenes.c2_cornell()
    w = h = a.size
    s = dev.default_stream()
    ext = torch.cuda.ExternalStream(s.cuda_stream())

    def timed(fn):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(ext); fn(); e1.record(ext); s.synchronize()
        return e0.elapsed_time(e1)

    pt = ex.PathTracer(dev, desc.meshes, w, h)
    pt.dispatch(a.spp, a.depth, count_rays=False); s.synchronize()   # warm-up
    pt.image.view().copy_from(np.zeros((w * h, 4), np.float32)); pt.seeds.view().copy_from(ex.seed_image(w, h)); pt.rays = [0, 0]
    ms = timed(lambda: [pt.dispatch(a.spp, a.depth, count_rays=False) for _ in range(a.dispatches)])
    hand_img, _ = pt.download()
    # ray counts from a counted replay on the same seeds
    pt.image.view().copy_from(np.zeros((w * h, 4), np.float32)); pt.seeds.view().copy_from(ex.seed_image(w, h))
    for _ in range(a.dispatches):
        pt.dispatch(a.spp, a.depth, count_rays=True)
    rays = sum(pt.rays)
    print(json.dumps({"variant": "hand", "size": w, "spp": a.spp * a.dispatches, "depth": a.depth, "ms_per_dispatch": ms / a.dispatches, "rays": rays,
                      "mrays_per_s": rays / ms / 1e3}))

    n = len(desc.meshes)
    vheap, iheap = dev.create_bindless_array(n), dev.create_bindless_array(n)
    for i, (vb, ib) in enumerate(zip(pt.vbuffers, pt.ibuffers)):
        vheap.emplace_buffer_async(i, vb); iheap.emplace_buffer_async(i, ib)
    s.submit([vheap.update_async(), iheap.update_async()])
    for poly in (True, False):
        image = dev.create_tex2d("Rgba32f", w, h); seeds = dev.create_tex2d("R32Uint", w, h)
        k = examples_ir.path_tracer_kernel(vheap.handle.id, iheap.handle.id, a.spp, a.depth, polynomial_sincos=poly)
        t0 = time.perf_counter()
        sh = dev.create_shader(C.addressof(k.km), keep=k)
        compile_s = time.perf_counter() - t0
        res = np.array([w, h], np.uint32)
        seeds.copy_from(ex.seed_image(w, h).reshape(h, w))
        sh.dispatch((w, h), image, seeds, pt.accel, res)   # warm-up
        image.copy_from(np.zeros((h, w, 4), np.float32)); seeds.copy_from(ex.seed_image(w, h).reshape(h, w))
        ms = timed(lambda: s.submit([sh.dispatch_async((w, h), image, seeds, pt.accel, res) for _ in range(a.dispatches)]))
        img = image.to_numpy()
        out = {"variant": "lowered" + ("_poly" if poly else "_libdevice"), "size": w, "spp": a.spp * a.dispatches, "depth": a.depth, "ms_per_dispatch": ms / a.dispatches,
               "create_shader_s": compile_s, "mrays_per_s": rays / ms / 1e3 if poly else None}
        if poly:
            out["bit_identical_to_hand"] = bool(np.array_equal(img.view(np.uint32), hand_img.view(np.uint32)))
        else:
            m0 = float((hand_img[..., :3] / hand_img[..., 3:4]).mean()); m1 = float((img[..., :3] / img[..., 3:4]).mean())
            out["mean_radiance_hand"] = m0; out["mean_radiance_lowered"] = m1
        print(json.dumps(out))
        sh.destroy(); image.destroy(); seeds.destroy()
    if a.cutout:
        # examples/path_tracer_cutout.rs: every ray a RayQuery with the stripes filter inlined as the candidate callback, instances non-opaque
        ptc = ex.PathTracer(dev, desc.meshes, w, h, opaque=False)
        vh2, ih2 = dev.create_bindless_array(n), dev.create_bindless_array(n)
        for i, (vb, ib) in enumerate(zip(ptc.vbuffers, ptc.ibuffers)):
            vh2.emplace_buffer_async(i, vb); ih2.emplace_buffer_async(i, ib)
        s.submit([vh2.update_async(), ih2.update_async()])
        image = dev.create_tex2d("Rgba32f", w, h); seeds = dev.create_tex2d("R32Uint", w, h)
        k = examples_ir.path_tracer_kernel(vh2.handle.id, ih2.handle.id, a.spp, a.depth, polynomial_sincos=True, cutout=True)
        t0 = time.perf_counter()
        sh = dev.create_shader(C.addressof(k.km), keep=k)
        compile_s = time.perf_counter() - t0
        res = np.array([w, h], np.uint32)
        seeds.copy_from(ex.seed_image(w, h).reshape(h, w))
        sh.dispatch((w, h), image, seeds, ptc.accel, res)   # warm-up
        image.copy_from(np.zeros((h, w, 4), np.float32)); seeds.copy_from(ex.seed_image(w, h).reshape(h, w))
        ms = timed(lambda: s.submit([sh.dispatch_async((w, h), image, seeds, ptc.accel, res) for _ in range(a.dispatches)]))
        print(json.dumps({"variant": "lowered_cutout_ray_query", "size": w, "spp": a.spp * a.dispatches, "depth": a.depth, "ms_per_dispatch": ms / a.dispatches,
                          "create_shader_s": compile_s, "note": "path_tracer_cutout.rs; rays not counted (paths differ from the opaque scene)"}))
        sh.destroy(); image.destroy(); seeds.destroy(); vh2.destroy(); ih2.destroy(); ptc.destroy()
    pt.destroy(); dev.close()


if __name__ == "__main__":
    main()
