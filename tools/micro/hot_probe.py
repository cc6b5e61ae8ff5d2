"""How many node visits of the C3 workload fall on the first N nodes of the BLAS (BFS order)?  Variant libraries built with
-DLCB_COUNT_HOT=N count them in TraceCounters.instance_entries (development probe)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "..")); sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "..", "tests"))
import luisa_compute_rs_b200 as lc
import scenes
dev = lc.Context().create_device("b200")
verts, tris = scenes.random_soup(1_000_000, 0x5EED0001)
vb, ib = dev.create_buffer_from_array(verts), dev.create_buffer_from_array(tris)
mesh = dev.create_mesh(vb.view(), ib.view()); mesh.build()
accel = dev.create_accel(); accel.push_mesh(mesh); accel.build()
n = 1 << 22
rays = scenes.incoherent_rays(n, seed=0x5EED0002)
rb, hb = dev.create_buffer(n, 32, 16), dev.create_buffer(n, 24, 8)
rb.view().copy_from(rays)
ctr = accel.intersect_counted(rb, hb, n, 0xFF)
print(os.environ.get("LC_B200_LIB"), {k: v / n for k, v in ctr.items()}, mesh.stats())
