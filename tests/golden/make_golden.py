"""Generates the committed fixtures in tests/golden/.

c1_closed_form_64.npz   analytic: the raytracing.rs triangle is in the plane z = 0 and rays start at (0,0,-1), so
                        t = 1/d.z and the hit point is (d.x/d.z, d.y/d.z, 0); barycentrics from the 2-D edge functions.
                        All in float64 from the float32 ray directions.  Independent of the oracle.
cornell_primary_48.npz, soup2k_rays4k.npz
                        regression vectors produced BY THE ORACLE's brute-force mode (the reference has no golden hit
                        buffers — SURVEY.md §4 — so these pin the oracle against drift, not against Embree).
Run:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as ol  # noqa: E402
import scenes  # noqa: E402


def c1_closed_form(w, h):
    rays = scenes.c1_rays(w, h)
    d = rays["dir"].astype(np.float64)
    t = 1.0 / d[:, 2]
    px, py = d[:, 0] * t, d[:, 1] * t
    v0, v1, v2 = np.array([-0.5, -0.5]), np.array([0.5, 0.0]), np.array([0.0, 0.5])
    def edge(a, b): return (b[0] - a[0]) * (py - a[1]) - (b[1] - a[1]) * (px - a[0])
    area = (v1[0] - v0[0]) * (v2[1] - v0[1]) - (v1[1] - v0[1]) * (v2[0] - v0[0])
    w0, w1, w2 = edge(v1, v2) / area, edge(v2, v0) / area, edge(v0, v1) / area
    hit = (w0 >= 0) & (w1 >= 0) & (w2 >= 0)
    return dict(hit=hit, t=t.astype(np.float32), bary=np.stack([w1, w2], 1).astype(np.float32),
                edge_margin=np.minimum(np.minimum(np.abs(w0), np.abs(w1)), np.abs(w2)).astype(np.float32))


def regression(desc, rays):
    h = ol.scene_from_desc(desc).trace_closest(rays, mode=ol.BRUTE)
    return dict(inst=h["inst"], prim=h["prim"], t_bits=h["committed_ray_t"].view(np.uint32), bary_bits=h["bary"].view(np.uint32))


if __name__ == "__main__":
    np.savez_compressed(os.path.join(HERE, "c1_closed_form_64.npz"), **c1_closed_form(64, 64))
    np.savez_compressed(os.path.join(HERE, "cornell_primary_48.npz"), **regression(scenes.c2_cornell(), scenes.c2_primary_rays(48, 48)))
    np.savez_compressed(os.path.join(HERE, "soup2k_rays4k.npz"), **regression(scenes.c3_soup(2000), scenes.incoherent_rays(4096)))
    print("golden fixtures written")
