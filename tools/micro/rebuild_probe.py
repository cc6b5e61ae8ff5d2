#!/usr/bin/env python
"""Where does a ForceBuild after PreferUpdate refits spend its time?  Wall clock vs device time per MeshBuild on a C4-sized terrain."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import luisa_compute_rs_b200 as lc
import scenes
nx = int(sys.argv[1]) if len(sys.argv) > 1 else 3164
dev = lc.Context().create_device("b200")
verts, tris = scenes.terrain(nx)
vb = dev.create_buffer_from_array(verts); ib = dev.create_buffer_from_array(tris)
mesh = dev.create_mesh(vb.view(), ib.view(), lc.AccelOption(allow_update=True))
def build(req, tag):
    t0 = time.perf_counter(); mesh.build(req); dt = (time.perf_counter() - t0) * 1e3
    print(f"{tag:28s} wall {dt:9.2f} ms   device {mesh.stats()['build_ms']:9.3f} ms   refit={mesh.stats()['was_refit']}", flush=True)
for i in range(3):
    build(lc.AccelBuildRequest.FORCE_BUILD, f"force #{i}")
for i in range(2):
    build(lc.AccelBuildRequest.PREFER_UPDATE, f"refit #{i}")
for i in range(3):
    build(lc.AccelBuildRequest.FORCE_BUILD, f"force after refit #{i}")
big = dev.create_buffer(8_000_000, 32, 16)
for i in range(2):
    build(lc.AccelBuildRequest.FORCE_BUILD, f"force after cudaMalloc #{i}")
fv = verts.copy(); fv[:, 1] += np.float32(0.01) * np.sin(np.float32(40.0) * verts[:, 0]).astype(np.float32)
vb.view().copy_from(fv)
for i in range(2):
    build(lc.AccelBuildRequest.FORCE_BUILD, f"force on displaced verts #{i}")
dev.close()
