#!/usr/bin/env python
"""One ForceBuild of an nx x nx terrain (default C4 size) — the target of ncu captures of the build kernels."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import luisa_compute_rs_b200 as lc
import scenes
nx = int(sys.argv[1]) if len(sys.argv) > 1 else 3164
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
dev = lc.Context().create_device("b200")
verts, tris = scenes.terrain(nx) if nx > 0 else scenes.random_soup(-nx, 0x5EED0001)
vb = dev.create_buffer_from_array(verts); ib = dev.create_buffer_from_array(tris)
mesh = dev.create_mesh(vb.view(), ib.view(), lc.AccelOption())
for _ in range(reps):
    mesh.build(lc.AccelBuildRequest.FORCE_BUILD)
print(mesh.stats())
dev.close()
