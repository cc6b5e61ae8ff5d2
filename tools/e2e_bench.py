#!/usr/bin/env python
"""Times lc_b200_trace_closest_host (pinned host rays in, pinned host hits out) on C3 for a given LC_B200_HOST_CHUNK."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np, torch
import luisa_compute_rs_b200 as lc, scenes
dev = lc.Context().create_device("b200")
verts, tris = scenes.random_soup(1_000_000, 0x5EED0001)
vb = dev.create_buffer_from_array(verts); ib = dev.create_buffer_from_array(tris)
mesh = dev.create_mesh(vb.view(), ib.view()); mesh.build()
accel = dev.create_accel(); accel.push_mesh(mesh); accel.build()
n = 1 << 24
rays = torch.from_numpy(scenes.incoherent_rays(n).view(np.uint8).reshape(-1)).pin_memory()
hits = torch.empty(n * 24, dtype=torch.uint8).pin_memory()
for _ in range(2):
    accel.intersect_host_ptr(rays.data_ptr(), hits.data_ptr(), n)
best = 1e9
for _ in range(5):
    t0 = time.perf_counter(); accel.intersect_host_ptr(rays.data_ptr(), hits.data_ptr(), n); best = min(best, time.perf_counter() - t0)
print(f"chunk={os.environ.get('LC_B200_HOST_CHUNK', 'default')}: {best * 1e3:.2f} ms  {n / best / 1e6:.1f} Mrays/s", flush=True)
dev.close()
