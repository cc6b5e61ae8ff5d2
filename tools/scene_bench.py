#!/usr/bin/env python
"""Full-size runs of BASELINE.json's configs C4 (20M-triangle terrain, per-frame refit + periodic rebuild, 4K primary and
secondary rays) and C5 (10 instances of a 5M-triangle mesh, 4K rays) on one GPU: build / refit ms, Mrays/s, and
size-independent checks (any-hit == closest-hit found, determinism, optional oracle sample).  One JSON line per config."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402


def camera_rays(w, h, origin, look_at, fov_deg=45.0):
    o = np.asarray(origin, np.float32)
    f = np.asarray(look_at, np.float32) - o; f /= np.linalg.norm(f)
    r = np.cross(f, np.float32([0, 1, 0])); r /= np.linalg.norm(r)
    u = np.cross(r, f)
    t = np.float32(np.tan(np.radians(fov_deg) / 2))
    x = ((np.arange(w, dtype=np.float32) + 0.5) / w * 2 - 1) * t * np.float32(w / h)
    y = (1 - (np.arange(h, dtype=np.float32) + 0.5) / h * 2) * t
    d = f[None, None, :] + x[None, :, None] * r[None, None, :] + y[:, None, None] * u[None, None, :]
    d = d.reshape(-1, 3).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    import scenes
    return scenes.make_rays(np.broadcast_to(o, d.shape), d, np.float32(1e-4), np.float32(1e30))


def secondary_rays(rays, hits, seed=5):
    """one cosine-ish hemisphere bounce per hit (around +y, the terrain's dominant normal), offset along it"""
    import luisa_compute_rs_b200 as lc
    rng = np.random.default_rng(seed)
    v = hits["inst"] != 0xFFFFFFFF
    p = rays["orig"][v] + rays["dir"][v] * hits["committed_ray_t"][v][:, None]
    d = rng.normal(size=p.shape).astype(np.float32); d[:, 1] = np.abs(d[:, 1]) + 0.1
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    n = np.zeros_like(p); n[:, 1] = 1
    import scenes
    return scenes.make_rays(lc.offset_ray_origin(p, n), d, np.float32(0.0), np.float32(1e30))


def timed(stream, fn, reps):
    import torch
    ext = torch.cuda.ExternalStream(stream.cuda_stream())
    fn(); stream.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    ev[0].record(ext)
    for k in range(reps):
        fn(); ev[k + 1].record(ext)
    stream.synchronize()
    return min(ev[k].elapsed_time(ev[k + 1]) for k in range(reps))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c4", choices=["c4", "c5"])
    ap.add_argument("--nx", type=int, default=0, help="terrain vertices per side (default: the config's full size)")
    ap.add_argument("--check", type=int, default=0, help="oracle sample size (builds the CPU BVH: slow at full size)")
    ap.add_argument("--frames", type=int, default=4)
    args = ap.parse_args()
    import luisa_compute_rs_b200 as lc
    import scenes
    dev = lc.Context().create_device("b200")
    stream = dev.create_stream()
    nx = args.nx or (3164 if args.config == "c4" else 1582)
    t0 = time.time()
    verts, tris = scenes.terrain(nx)
    gen_s = time.time() - t0
    out = {"config": args.config, "nx": nx, "triangles_per_mesh": int(tris.shape[0]), "host_scene_generation_s": round(gen_s, 2)}
    vb = dev.create_buffer_from_array(verts); ib = dev.create_buffer_from_array(tris)
    mesh = dev.create_mesh(vb.view(), ib.view(), lc.AccelOption(allow_update=True))
    builds = []
    for _ in range(3):
        mesh.build(lc.AccelBuildRequest.FORCE_BUILD); builds.append(mesh.stats()["build_ms"])
    st = mesh.stats()
    out.update(blas_build_ms=[round(b, 3) for b in builds], wide_nodes=int(st["wide_node_count"]), bvh_bytes=int(st["bvh_bytes"]), depth=int(st["max_depth"]))
    accel = dev.create_accel(lc.AccelOption(allow_update=True))
    if args.config == "c4":
        accel.push_mesh(mesh)
        cam_o, cam_at = [0.5, 0.6, -0.6], [0.5, 0.0, 0.5]
    else:
        for k in range(10):  # 5 x 2 grid, yaw 36 deg * k
            t = np.eye(4, dtype=np.float32); t[:3, :] = scenes.rotation_y(36.0 * k)
            t[:3, 3] = [1.2 * (k % 5), 0.0, 1.2 * (k // 5)]
            accel.push_mesh(mesh, t)
        cam_o, cam_at = [3.0, 2.5, -3.0], [3.0, 0.0, 1.0]
    accel.build()
    out["tlas_build_ms"] = round(accel.stats()["build_ms"], 3)
    out["instances"] = len(accel.instance_handles)
    out["total_triangles"] = int(tris.shape[0]) * len(accel.instance_handles)
    # ---- 4K primary + secondary rays ----
    w, h = 3840, 2160
    prim = camera_rays(w, h, cam_o, cam_at)
    n = prim.shape[0]
    rb = dev.create_buffer_from_array(prim); hb = dev.create_buffer(n, 24, 8); ob = dev.create_buffer(n, 4, 4)
    ms_p = timed(stream, lambda: accel.intersect(rb, hb, n, 0xFF, stream), 5)
    hits = hb.view().to_numpy(lc.SurfaceHit)
    hits2 = hits.copy(); accel.intersect(rb, hb, n, 0xFF, stream); stream.synchronize(); hb.view().copy_to(hits2)
    ms_a = timed(stream, lambda: accel.intersect_any(rb, ob, n, 0xFF, stream), 5)
    occ = ob.view().to_numpy(np.uint32)
    found = hits["inst"] != lc.INVALID
    sec = secondary_rays(prim, hits)
    ns = sec.shape[0]
    rb2 = dev.create_buffer_from_array(sec); hb2 = dev.create_buffer(ns, 24, 8)
    ms_s = timed(stream, lambda: accel.intersect(rb2, hb2, ns, 0xFF, stream), 5)
    hits_s = hb2.view().to_numpy(lc.SurfaceHit)
    out.update(primary_rays=n, primary_ms=round(ms_p, 3), primary_mrays_s=round(n / ms_p / 1e3, 1), primary_hit_rate=round(float(found.mean()), 4),
               any_ms=round(ms_a, 3), any_mrays_s=round(n / ms_a / 1e3, 1), secondary_rays=int(ns), secondary_ms=round(ms_s, 3),
               secondary_mrays_s=round(ns / ms_s / 1e3, 1), secondary_hit_rate=round(float((hits_s["inst"] != lc.INVALID).mean()), 4),
               check_any_equals_closest=bool(np.array_equal(occ != 0, found)), check_deterministic=bool(hits.tobytes() == hits2.tobytes()))
    if args.check:
        import oracle_lib as ol
        desc = scenes.SceneDesc(); mid = desc.add_mesh(verts, tris)
        if args.config == "c4":
            desc.add_instance(mid)
        else:
            for k in range(10):
                t = scenes.rotation_y(36.0 * k); t[:, 3] = [1.2 * (k % 5), 0.0, 1.2 * (k // 5)]
                desc.add_instance(mid, t)
        t0 = time.time(); o = ol.scene_from_desc(desc); out["oracle_build_s"] = round(time.time() - t0, 1)
        pick = np.random.default_rng(3).integers(0, n, args.check)
        out["check_oracle_primary_identical"] = bool(hits[pick].tobytes() == o.trace_closest(prim[pick]).tobytes())
        pick2 = np.random.default_rng(4).integers(0, ns, args.check)
        out["check_oracle_secondary_identical"] = bool(hits_s[pick2].tobytes() == o.trace_closest(sec[pick2]).tobytes())
        o.close()
    # ---- per-frame update: refit every frame, rebuild at the end (C4) ----
    if args.config == "c4":
        refit_ms, tlas_ms, trace_ms = [], [], []
        x = verts[:, 0]
        for f in range(1, args.frames + 1):
            fv = verts.copy(); fv[:, 1] += np.float32(0.01) * np.sin(np.float32(40.0) * x + np.float32(0.3 * f)).astype(np.float32)
            vb.view().copy_from(fv)
            mesh.build(lc.AccelBuildRequest.PREFER_UPDATE); refit_ms.append(mesh.stats()["build_ms"]); assert mesh.stats()["was_refit"] == 1
            accel.build(lc.AccelBuildRequest.PREFER_UPDATE); tlas_ms.append(accel.stats()["build_ms"])
            trace_ms.append(timed(stream, lambda: accel.intersect(rb, hb, n, 0xFF, stream), 2))
        mesh.build(lc.AccelBuildRequest.FORCE_BUILD); rebuild_ms = mesh.stats()["build_ms"]
        accel.build()
        ms_after = timed(stream, lambda: accel.intersect(rb, hb, n, 0xFF, stream), 2)
        out.update(refit_ms=[round(r, 3) for r in refit_ms], refit_tlas_ms=[round(r, 3) for r in tlas_ms], primary_ms_after_refits=[round(r, 3) for r in trace_ms],
                   rebuild_ms=round(rebuild_ms, 3), primary_ms_after_rebuild=round(ms_after, 3))
        nv = verts.shape[0]
        peak = 6547.8
        refit_bytes = 12 * nv + 12 * tris.shape[0] + 2 * st["wide_node_count"] * 128 + tris.shape[0] * 64
        build_bytes = 12 * tris.shape[0] + 12 * nv + st["bvh_bytes"]
        out.update(build_roofline_frac=round(build_bytes / (min(builds) * 1e-3) / 1e9 / peak, 4), refit_roofline_frac=round(refit_bytes / (min(refit_ms) * 1e-3) / 1e9 / peak, 4))
    print(json.dumps(out), flush=True)
    dev.close()


if __name__ == "__main__":
    main()
