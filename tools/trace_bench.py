#!/usr/bin/env python
"""Quick kernel-level timing of trace_closest / trace_any / BLAS build on the C3 workload (development aid for variant
sweeps; bench.py is the contract).  Prints one line per run; LC_B200_LIB selects the library variant."""
import argparse
import hashlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tris", type=int, default=1_000_000)
    ap.add_argument("--rays", type=int, default=1 << 24)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--check", type=int, default=0, help="compare this many sampled rays with the oracle")
    ap.add_argument("--scene", default="soup", choices=["soup", "terrain"])
    ap.add_argument("--tag", default=os.environ.get("LC_B200_LIB", "default"))
    args = ap.parse_args()
    import torch
    import luisa_compute_rs_b200 as lc
    import scenes
    dev = lc.Context().create_device("b200")
    if args.scene == "soup":
        verts, tris = scenes.random_soup(args.tris, 0x5EED0001)
        rays = scenes.incoherent_rays(args.rays, seed=0x5EED0002)
    else:
        nx = int(np.sqrt(args.tris / 2)) + 1
        verts, tris = scenes.terrain(nx)
        rays = scenes.incoherent_rays(args.rays, seed=0x5EED0002)
        rays["orig"][:, 1] = rays["orig"][:, 1] * 0.3 + 0.1
    vb = dev.create_buffer_from_array(verts); ib = dev.create_buffer_from_array(tris)
    mesh = dev.create_mesh(vb.view(), ib.view(), lc.AccelOption())
    bms = []
    for _ in range(4):
        mesh.build(); bms.append(mesh.stats()["build_ms"])
    accel = dev.create_accel(); accel.push_mesh(mesh); accel.build()
    n = args.rays
    rb = dev.create_buffer_from_array(rays); hb = dev.create_buffer(n, 24, 8); ob = dev.create_buffer(n, 4, 4)
    stream = dev.create_stream()
    ext = torch.cuda.ExternalStream(stream.cuda_stream())
    def timed(fn):
        for _ in range(2):
            fn()
        stream.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.reps + 1)]
        ev[0].record(ext)
        for k in range(args.reps):
            fn(); ev[k + 1].record(ext)
        stream.synchronize()
        return min(ev[k].elapsed_time(ev[k + 1]) for k in range(args.reps))
    t_c = timed(lambda: accel.intersect(rb, hb, n, 0xFF, stream))
    hits = hb.view().to_numpy(lc.SurfaceHit)
    t_a = timed(lambda: accel.intersect_any(rb, ob, n, 0xFF, stream))
    digest = hashlib.sha1(hits.tobytes()).hexdigest()[:12]
    msg = (f"[{args.tag}] {args.scene} tris={tris.shape[0]} rays={n}: closest {t_c:.3f} ms = {n / t_c / 1e3:.1f} Mrays/s | any {t_a:.3f} ms = "
           f"{n / t_a / 1e3:.1f} Mrays/s | build ms {['%.2f' % b for b in bms]} nodes={mesh.stats()['wide_node_count']} | hits sha1 {digest}")
    if args.check:
        import oracle_lib as ol
        desc = scenes.SceneDesc(); desc.add_instance(desc.add_mesh(verts, tris))
        o = ol.scene_from_desc(desc)
        pick = np.random.default_rng(1).integers(0, n, args.check)
        want = o.trace_closest(rays[pick])
        ok = hits[pick].tobytes() == want.tobytes()
        msg += f" | oracle sample {args.check}: {'IDENTICAL' if ok else 'MISMATCH'}"
    print(msg, flush=True)
    dev.close()


if __name__ == "__main__":
    main()
