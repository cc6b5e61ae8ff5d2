"""ctypes binding of oracle/liboracle.so — the checker.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module."""
import ctypes as C
import os
import subprocess

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_SO = os.path.join(_ROOT, "oracle", "liboracle.so")
_lib = None


class Filter(C.Structure):
    _fields_ = [("kind", C.c_int), ("radius", C.c_float), ("bits", C.c_void_p), ("first_bit", C.c_void_p), ("stripe_freq", C.c_void_p), ("stripe_keep", C.c_void_p)]


class Mod(C.Structure):
    _fields_ = [("index", C.c_uint32), ("user_id", C.c_uint32), ("flags", C.c_uint32), ("visibility", C.c_uint32),
                ("mesh", C.c_uint64), ("affine", C.c_float * 12)]


def build():
    subprocess.run(["make", "-C", os.path.join(_ROOT, "oracle")], check=True, capture_output=True)


def lib():
    global _lib
    if _lib is None:
        build()  # make: a no-op when liboracle.so is newer than oracle.c, and never a stale checker after an edit
        L = C.CDLL(_SO)
        L.oracle_scene_new.restype = C.c_void_p
        L.oracle_scene_free.argtypes = [C.c_void_p]
        L.oracle_mesh_new.argtypes = [C.c_void_p]
        L.oracle_mesh_new.restype = C.c_uint64
        L.oracle_mesh_set.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t]
        L.oracle_mesh_commit.argtypes = [C.c_void_p, C.c_uint64]
        L.oracle_accel_update.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(Mod), C.c_size_t]
        L.oracle_instance_transform.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(C.c_float)]
        L.oracle_instance_user_id.argtypes = [C.c_void_p, C.c_uint32]
        L.oracle_instance_user_id.restype = C.c_uint32
        L.oracle_instance_visibility.argtypes = [C.c_void_p, C.c_uint32]
        L.oracle_instance_visibility.restype = C.c_uint32
        L.oracle_trace_closest.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_void_p, C.c_int, C.c_int]
        L.oracle_trace_any.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_void_p, C.c_int, C.c_int]
        L.oracle_trace_closest_f64.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.oracle_ray_query.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_int, C.POINTER(Filter), C.c_void_p, C.c_int, C.c_int]
        L.oracle_path_tracer.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_float,
                                         C.c_int, C.POINTER(C.c_uint64)]
        L.oracle_offset_ray_origin.argtypes = [C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float)]
        L.oracle_canonical_triangle.argtypes = [C.POINTER(C.c_float)] * 2 + [C.c_float, C.c_float] + [C.POINTER(C.c_float)] * 6
        L.oracle_canonical_triangle.restype = C.c_int
        L.oracle_invert_affine.argtypes = [C.POINTER(C.c_float), C.POINTER(C.c_float)]
        L.oracle_curve_set.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_size_t]
        L.oracle_canonical_cone.argtypes = [C.POINTER(C.c_float)] * 2 + [C.c_float, C.c_float] + [C.POINTER(C.c_float)] * 4
        L.oracle_canonical_cone.restype = C.c_int
        L.oracle_hw_threads.restype = C.c_int
        _lib = L
    return _lib


def canonical_cone(o, d, tmin, tmax, A, B):
    """oracle canon_cone on one ray and one rounded cone (A, B = (x, y, z, radius)): (hit, t, s)"""
    f = lambda a: (C.c_float * len(a))(*[float(x) for x in a])
    t, s = C.c_float(), C.c_float()
    hit = lib().oracle_canonical_cone(f(o), f(d), tmin, tmax, f(A), f(B), C.byref(t), C.byref(s))
    return bool(hit), t.value, s.value


HIT = np.dtype([("inst", "<u4"), ("prim", "<u4"), ("bary", "<f4", (2,)), ("committed_ray_t", "<f4"), ("_pad", "<u4")])
COMMITTED = np.dtype([("inst", "<u4"), ("prim", "<u4"), ("bary", "<f4", (2,)), ("hit_type", "<u4"), ("committed_ray_t", "<f4")])
BRUTE, BVH, WIDE = 0, 1, 2   # WIDE: the 8-wide AVX2 traversal (CPU baseline); same canonical hits


class OracleScene:
    def __init__(self):
        self.L = lib()
        self.s = C.c_void_p(self.L.oracle_scene_new())
        self.keep = []

    def add_mesh(self, verts, tris, stride=None):
        verts = np.ascontiguousarray(verts, dtype=np.float32)
        tris = np.ascontiguousarray(tris, dtype=np.uint32)
        self.keep += [verts, tris]
        m = self.L.oracle_mesh_new(self.s)
        self.L.oracle_mesh_set(self.s, m, verts.ctypes.data, stride or verts.strides[0], verts.shape[0], tris.ctypes.data, 12, tris.shape[0])
        self.L.oracle_mesh_commit(self.s, m)
        return m

    def add_curve(self, basis, cps, segs):
        """control points (n, 4) float32 {x, y, z, radius}; segs (m,) uint32 first-control-point indices"""
        cps = np.ascontiguousarray(cps, dtype=np.float32)
        segs = np.ascontiguousarray(segs, dtype=np.uint32)
        self.keep += [cps, segs]
        m = self.L.oracle_mesh_new(self.s)
        self.L.oracle_curve_set(self.s, m, int(basis), cps.ctypes.data, cps.strides[0], cps.shape[0], segs.ctypes.data, segs.shape[0])
        return m

    def commit_mesh(self, m):
        self.L.oracle_mesh_commit(self.s, m)

    def update(self, instance_count, mods):
        arr = (Mod * max(len(mods), 1))()
        for i, m in enumerate(mods):
            arr[i] = Mod(m["index"], m.get("user_id", 0), m["flags"], m.get("visibility", 0), m.get("mesh", 0),
                         (C.c_float * 12)(*np.asarray(m.get("affine", [1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0]), dtype=np.float32).reshape(12)))
        self.L.oracle_accel_update(self.s, instance_count, arr, len(mods))

    def trace_closest(self, rays, mask=0xFF, mode=BVH, threads=0):
        rays = np.ascontiguousarray(rays)
        hits = np.zeros(rays.shape[0], dtype=HIT)
        self.L.oracle_trace_closest(self.s, rays.ctypes.data, rays.shape[0], mask, hits.ctypes.data, mode, threads)
        return hits

    def trace_any(self, rays, mask=0xFF, mode=BVH, threads=0):
        rays = np.ascontiguousarray(rays)
        occ = np.zeros(rays.shape[0], dtype=np.uint32)
        self.L.oracle_trace_any(self.s, rays.ctypes.data, rays.shape[0], mask, occ.ctypes.data, mode, threads)
        return occ

    def ray_query(self, rays, mask=0xFF, terminate_on_first=False, kind=0, radius=0.0, bits=None, first_bit=None, mode=BVH, threads=0):
        rays = np.ascontiguousarray(rays)
        out = np.zeros(rays.shape[0], dtype=COMMITTED)
        f = Filter(kind, radius, None, None)
        if bits is not None:
            bits = np.ascontiguousarray(bits, dtype=np.uint32); first_bit = np.ascontiguousarray(first_bit, dtype=np.uint32)
            f.bits, f.first_bit = bits.ctypes.data, first_bit.ctypes.data
        self.L.oracle_ray_query(self.s, rays.ctypes.data, rays.shape[0], mask, int(terminate_on_first), C.byref(f), out.ctypes.data, mode, threads)
        return out

    def truth(self, rays, mask=0xFF, mode=BVH, threads=0):
        rays = np.ascontiguousarray(rays)
        hits = np.zeros(rays.shape[0], dtype=HIT)
        amb = np.zeros(rays.shape[0], dtype=np.uint8)
        self.L.oracle_trace_closest_f64(self.s, rays.ctypes.data, rays.shape[0], mask, hits.ctypes.data, amb.ctypes.data, mode, threads)
        return hits, amb

    def instance_transform(self, i):
        out = (C.c_float * 12)()
        self.L.oracle_instance_transform(self.s, i, out)
        return np.array(out, dtype=np.float32).reshape(3, 4)

    def instance_user_id(self, i):
        return self.L.oracle_instance_user_id(self.s, i)

    def instance_visibility(self, i):
        return self.L.oracle_instance_visibility(self.s, i)

    def close(self):
        if self.s:
            self.L.oracle_scene_free(self.s)
            self.s = None


def scene_from_desc(desc):
    """Build an OracleScene from tests.scenes.SceneDesc with the modification sequence Accel::push_mesh records."""
    o = OracleScene()
    ids = [o.add_mesh(v, t) for v, t in desc.meshes]
    mods = []
    for k, inst in enumerate(desc.instances):
        flags = 1 | 2 | 16 | 32 | (4 if inst["opaque"] else 8)
        mods.append(dict(index=k, user_id=inst["user_id"], flags=flags, visibility=inst["mask"], mesh=ids[inst["mesh"]], affine=inst["transform"].reshape(12)))
    o.update(len(desc.instances), mods)
    return o


def path_tracer_dispatch(oracle_scene, meshes, image, seeds, width, height, spp, max_depth, tan_half_fov, threads=0):
    """One dispatch of the CPU restatement of examples/path_tracer.rs; `image` (h,w,4 float32) and `seeds` (h*w uint32) are updated in place."""
    vs = [np.ascontiguousarray(v, np.float32) for v, _ in meshes]
    ts = [np.ascontiguousarray(t, np.uint32) for _, t in meshes]
    vh = (C.c_void_p * len(vs))(*[v.ctypes.data for v in vs])
    ih = (C.c_void_p * len(ts))(*[t.ctypes.data for t in ts])
    counts = (C.c_uint64 * 2)()
    lib().oracle_path_tracer(oracle_scene.s, vh, ih, image.ctypes.data, seeds.ctypes.data, width, height, spp, max_depth, float(tan_half_fov), threads, counts)
    return counts[0], counts[1]


def path_tracer_cutout_dispatch(oracle_scene, meshes, image, seeds, width, height, spp, max_depth, tan_half_fov, stripe_freq, stripe_keep, threads=0):
    """One dispatch of the CPU restatement of examples/path_tracer_cutout.rs: ray queries whose candidates pass the stripes filter
    (stripe_freq / stripe_keep: one float per instance, freq 0 = no cut-out)."""
    vs = [np.ascontiguousarray(v, np.float32) for v, _ in meshes]
    ts = [np.ascontiguousarray(t, np.uint32) for _, t in meshes]
    vh = (C.c_void_p * len(vs))(*[v.ctypes.data for v in vs])
    ih = (C.c_void_p * len(ts))(*[t.ctypes.data for t in ts])
    fr, kp = np.ascontiguousarray(stripe_freq, np.float32), np.ascontiguousarray(stripe_keep, np.float32)
    flt = Filter(4, 0.0, None, None, fr.ctypes.data, kp.ctypes.data)
    counts = (C.c_uint64 * 2)()
    L = lib()
    L.oracle_path_tracer_cutout.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_float,
                                            C.POINTER(Filter), C.c_int, C.c_void_p]
    L.oracle_path_tracer_cutout(oracle_scene.s, vh, ih, image.ctypes.data, seeds.ctypes.data, width, height, spp, max_depth, float(tan_half_fov), C.byref(flt), threads, counts)
    return counts[0], counts[1]


def offset_ray_origin(p, n):
    L = lib()
    p = np.ascontiguousarray(p, dtype=np.float32)
    n = np.ascontiguousarray(n, dtype=np.float32)
    out = np.empty_like(p)
    fp = C.POINTER(C.c_float)
    for i in range(p.shape[0]):
        L.oracle_offset_ray_origin(p[i].ctypes.data_as(fp), n[i].ctypes.data_as(fp), out[i].ctypes.data_as(fp))
    return out


def canonical_triangle(o, d, tmin, tmax, v0, v1, v2):
    L = lib()
    fp = C.POINTER(C.c_float)
    arrs = [np.ascontiguousarray(a, dtype=np.float32) for a in (o, d, v0, v1, v2)]
    t, u, v = C.c_float(), C.c_float(), C.c_float()
    ok = L.oracle_canonical_triangle(arrs[0].ctypes.data_as(fp), arrs[1].ctypes.data_as(fp), tmin, tmax, arrs[2].ctypes.data_as(fp),
                                     arrs[3].ctypes.data_as(fp), arrs[4].ctypes.data_as(fp), C.byref(t), C.byref(u), C.byref(v))
    return (bool(ok), t.value, u.value, v.value)
