"""Host side of luisa_compute/examples/path_tracer.rs for the B200 device: scene upload (one mesh + instance per OBJ group,
path_tracer.rs:213-249), the per-dispatch call of the hand-lowered kernel (csrc/path_tracer.cu) and the accumulation loop
(path_tracer.rs:537-558)."""
import ctypes as C

import numpy as np

from . import _abi as abi
from .rtx import AccelBuildRequest, AccelOption

SPP_PER_DISPATCH = 32   # path_tracer.rs:179
MAX_DEPTH = 10          # path_tracer.rs:358
TAN_HALF_FOV = np.float32(np.tan(np.float32(0.5) * (np.float32(27.8) * np.float32(np.pi) / np.float32(180.0))))  # path_tracer.rs:293-302


def seed_image(width, height, seed=0xC0FFEE):
    """The example seeds with thread_rng (path_tracer.rs:486-490); fixed here: low 32 bits of splitmix64(seed + pixel)."""
    z = np.arange(width * height, dtype=np.uint64) + np.uint64(seed) + np.uint64(0x9E3779B97F4A7C15)
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return ((z ^ (z >> np.uint64(31))) & np.uint64(0xFFFFFFFF)).astype(np.uint32)


class PathTracer:
    def __init__(self, device, meshes, width, height, seed=0xC0FFEE, opaque=True):
        """meshes: list of (vertices float32 (nv,3), triangles uint32 (nt,3)), one instance each with identity transform;
        opaque=False is path_tracer_cutout.rs:250 (candidates of every instance reach the ray-query callback)."""
        self.device, self.width, self.height = device, width, height
        self.vbuffers, self.ibuffers, self.meshes = [], [], []
        self.accel = device.create_accel(AccelOption())
        for verts, tris in meshes:
            vb = device.create_buffer_from_array(np.ascontiguousarray(verts, np.float32))
            ib = device.create_buffer_from_array(np.ascontiguousarray(tris, np.uint32))
            m = device.create_mesh(vb.view(), ib.view(), AccelOption())
            m.build(AccelBuildRequest.FORCE_BUILD)
            self.accel.push_mesh(m, np.eye(4, dtype=np.float32), 255, opaque)
            self.vbuffers.append(vb); self.ibuffers.append(ib); self.meshes.append(m)
        self.accel.build(AccelBuildRequest.FORCE_BUILD)
        self.image = device.create_buffer(width * height, 16, 16)       # Tex2d<Float4>, zero-initialised
        self.seeds = device.create_buffer_from_array(seed_image(width, height, seed))
        n = len(meshes)
        self._vh = (abi.Handle * n)(*[b.handle for b in self.vbuffers])
        self._ih = (abi.Handle * n)(*[b.handle for b in self.ibuffers])
        self.rays = [0, 0]

    def dispatch(self, spp_per_dispatch=SPP_PER_DISPATCH, max_depth=MAX_DEPTH, stream=None, count_rays=True):
        s = stream or self.device.default_stream()
        args = abi.PathTracerArgs(self.accel.handle, self._vh, self._ih, len(self.meshes), self.image.handle, self.seeds.handle, self.width, self.height,
                                  spp_per_dispatch, max_depth, float(TAN_HALF_FOV))
        counts = (C.c_uint64 * 2)()
        self.device.lib.lc_b200_example_path_tracer(self.device.handle, s.handle, C.byref(args), counts if count_rays else None)
        if count_rays:
            self.rays[0] += counts[0]; self.rays[1] += counts[1]

    def download(self):
        img = self.image.view().to_numpy(np.float32).reshape(self.height, self.width, 4)
        return img, self.seeds.view().to_numpy(np.uint32)

    def destroy(self):
        self.accel.destroy()
        for m in self.meshes:
            m.destroy()
        for b in self.vbuffers + self.ibuffers + [self.image, self.seeds]:
            b.destroy()
