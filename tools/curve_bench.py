"""Curves on one B200: CurveBuild time and closest-hit Mrays/s of the batch entry point on a field of random-walk strands.
    python tools/curve_bench.py [--strands 100000] [--points 13] [--basis 1] [--rays 4194304]
Development measurement (the curve path is §8f rank 3, not the headline metric); prints one JSON line."""
import argparse, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import luisa_compute_rs_b200 as lc


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--strands", type=int, default=100000)
    ap.add_argument("--points", type=int, default=13)
    ap.add_argument("--basis", type=int, default=1, help="0 linear, 1 B-spline, 2 Catmull-Rom, 3 Bezier")
    ap.add_argument("--rays", type=int, default=1 << 22)
    a = ap.parse_args()
    rng = np.random.default_rng(1)
    ns, pp = a.strands, a.points
    p = np.zeros((ns, pp, 4), np.float32)
    p[:, 0, :3] = rng.random((ns, 3), dtype=np.float32)
    d = rng.normal(size=(ns, 3)).astype(np.float32); d /= np.linalg.norm(d, axis=1, keepdims=True)
    for j in range(1, pp):
        d = d + np.float32(0.5) * rng.normal(size=(ns, 3)).astype(np.float32); d /= np.linalg.norm(d, axis=1, keepdims=True)
        p[:, j, :3] = p[:, j - 1, :3] + np.float32(0.004) * d
    p[:, :, 3] = rng.uniform(0.0002, 0.0006, (ns, pp)).astype(np.float32)
    per = {0: range(pp - 1), 3: range(0, pp - 3, 3)}.get(a.basis, range(pp - 3))
    segs = (np.arange(ns, dtype=np.uint32)[:, None] * pp + np.asarray(list(per), np.uint32)[None, :]).reshape(-1)
    dev = lc.Context().create_device("b200")
    cpb, sgb = dev.create_buffer_from_array(p.reshape(-1, 4)), dev.create_buffer_from_array(segs)
    curve = dev.create_curve(a.basis, cpb.view(), sgb.view())
    builds = []
    for _ in range(4):
        curve.build(); builds.append(curve_stats(dev, curve)["build_ms"])
    accel = dev.create_accel(); accel.push_curve(curve); accel.build()
    n = a.rays
    rays = np.zeros((n, 8), np.float32)
    rays[:, 0:3] = rng.random((n, 3), dtype=np.float32)
    dd = rng.normal(size=(n, 3)).astype(np.float32); rays[:, 4:7] = dd / np.linalg.norm(dd, axis=1, keepdims=True)
    rays[:, 3] = 1e-4; rays[:, 7] = 1e30
    rb, hb = dev.create_buffer(n, 32, 16), dev.create_buffer(n, 24, 8)
    rb.view().copy_from(rays)
    s = dev.default_stream()
    for _ in range(2):
        accel.intersect(rb, hb, n); s.synchronize()
    t0 = time.perf_counter()
    reps = 5
    for _ in range(reps):
        accel.intersect(rb, hb, n)
    s.synchronize()
    dt = (time.perf_counter() - t0) / reps
    hits = np.zeros(n, dtype=lc.SurfaceHit); hb.view().copy_to(hits)
    st = curve_stats(dev, curve)
    print(json.dumps({"config": "curve_bench", "basis": a.basis, "segments": int(segs.shape[0]), "pieces": int(st["primitive_count"]),
                      "build_ms": float(min(builds)), "wide_nodes": int(st["wide_node_count"]), "rays": n, "trace_ms": dt * 1e3,
                      "mrays_per_s": n / dt / 1e6, "hit_rate": float((hits["inst"] != lc.INVALID).mean())}))


def curve_stats(dev, curve):
    import ctypes as C
    st = lc._abi.BuildStats()
    dev.lib.lc_b200_mesh_stats(dev.handle, curve.handle, C.byref(st))
    return st.as_dict()


if __name__ == "__main__":
    main()
