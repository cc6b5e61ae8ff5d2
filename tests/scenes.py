"""Deterministic scenes and ray sets for the parity tests and bench.py (SURVEY.md §8d "concrete synthetic inputs").

Everything here is numpy-only input generation; it computes no intersections.
"""
import numpy as np

RAY = np.dtype([("orig", "<f4", (3,)), ("tmin", "<f4"), ("dir", "<f4", (3,)), ("tmax", "<f4")])
HIT = np.dtype([("inst", "<u4"), ("prim", "<u4"), ("bary", "<f4", (2,)), ("committed_ray_t", "<f4"), ("_pad", "<u4")])
IDENTITY34 = np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 0]], dtype=np.float32)


class SceneDesc:
    """meshes: list of (vertices float32 (nv,3), triangles uint32 (nt,3));
    instances: list of dicts {mesh, transform(3x4), mask, opaque, user_id}."""

    def __init__(self):
        self.meshes = []
        self.instances = []

    def add_mesh(self, verts, tris):
        self.meshes.append((np.ascontiguousarray(verts, dtype=np.float32), np.ascontiguousarray(tris, dtype=np.uint32).reshape(-1, 3)))
        return len(self.meshes) - 1

    def add_instance(self, mesh, transform=None, mask=0xFF, opaque=True, user_id=0):
        t = IDENTITY34 if transform is None else np.asarray(transform, dtype=np.float32)[:3, :]
        self.instances.append(dict(mesh=mesh, transform=np.ascontiguousarray(t), mask=mask, opaque=opaque, user_id=user_id))
        return len(self.instances) - 1

    def triangle_count(self):
        return sum(self.meshes[i["mesh"]][1].shape[0] for i in self.instances)


def make_rays(o, d, tmin, tmax):
    o = np.asarray(o, dtype=np.float32)
    r = np.empty(o.shape[0], dtype=RAY)
    r["orig"] = o
    r["dir"] = np.asarray(d, dtype=np.float32)
    r["tmin"] = tmin
    r["tmax"] = tmax
    return r


# ---- C1: examples/raytracing.rs:32-63 ---------------------------------------------------------
def c1_triangle():
    s = SceneDesc()
    m = s.add_mesh([[-0.5, -0.5, 0.0], [0.5, 0.0, 0.0], [0.0, 0.5, 0.0]], [[0, 1, 2]])
    s.add_instance(m)
    return s


def c1_rays(w=1024, h=1024):
    """o = (0,0,-1), d = normalize((2x/W-1, 2y/H-1, 0) - o), tmin 1e-3, tmax 1e9 (raytracing.rs:44-56), fp32 throughout."""
    x, y = np.meshgrid(np.arange(w, dtype=np.float32), np.arange(h, dtype=np.float32))
    xy = np.stack([x / np.float32(w), y / np.float32(h)], -1).reshape(-1, 2)
    xy = np.float32(2.0) * xy - np.float32(1.0)
    o = np.zeros((w * h, 3), np.float32)
    o[:, 2] = -1.0
    d = np.concatenate([xy, np.zeros((w * h, 1), np.float32)], 1) - o
    d = (d / np.sqrt((d * d).sum(1, keepdims=True, dtype=np.float32))).astype(np.float32)
    return make_rays(o, d, np.float32(1e-3), np.float32(1e9))


# ---- C2: the Cornell box of examples/path_tracer.rs:35-177 (public-domain data set by Cardenas & McGuire) ----
# one quad = 4 vertices; tobj triangulates a quad (a,b,c,d) as (a,b,c),(a,c,d); one mesh + one instance per
# OBJ group in file order (floor, ceiling, backWall, rightWall, leftWall, shortBox, tallBox, light = inst 7).
_CBOX = {
    "floor": [[(-1.01, 0.00, 0.99), (1.00, 0.00, 0.99), (1.00, 0.00, -1.04), (-0.99, 0.00, -1.04)]],
    "ceiling": [[(-1.02, 1.99, 0.99), (-1.02, 1.99, -1.04), (1.00, 1.99, -1.04), (1.00, 1.99, 0.99)]],
    "backWall": [[(-0.99, 0.00, -1.04), (1.00, 0.00, -1.04), (1.00, 1.99, -1.04), (-1.02, 1.99, -1.04)]],
    "rightWall": [[(1.00, 0.00, -1.04), (1.00, 0.00, 0.99), (1.00, 1.99, 0.99), (1.00, 1.99, -1.04)]],
    "leftWall": [[(-1.01, 0.00, 0.99), (-0.99, 0.00, -1.04), (-1.02, 1.99, -1.04), (-1.02, 1.99, 0.99)]],
    "shortBox": [
        [(0.53, 0.60, 0.75), (0.70, 0.60, 0.17), (0.13, 0.60, 0.00), (-0.05, 0.60, 0.57)],
        [(-0.05, 0.00, 0.57), (-0.05, 0.60, 0.57), (0.13, 0.60, 0.00), (0.13, 0.00, 0.00)],
        [(0.53, 0.00, 0.75), (0.53, 0.60, 0.75), (-0.05, 0.60, 0.57), (-0.05, 0.00, 0.57)],
        [(0.70, 0.00, 0.17), (0.70, 0.60, 0.17), (0.53, 0.60, 0.75), (0.53, 0.00, 0.75)],
        [(0.13, 0.00, 0.00), (0.13, 0.60, 0.00), (0.70, 0.60, 0.17), (0.70, 0.00, 0.17)],
        [(0.53, 0.00, 0.75), (0.70, 0.00, 0.17), (0.13, 0.00, 0.00), (-0.05, 0.00, 0.57)]],
    "tallBox": [
        [(-0.53, 1.20, 0.09), (0.04, 1.20, -0.09), (-0.14, 1.20, -0.67), (-0.71, 1.20, -0.49)],
        [(-0.53, 0.00, 0.09), (-0.53, 1.20, 0.09), (-0.71, 1.20, -0.49), (-0.71, 0.00, -0.49)],
        [(-0.71, 0.00, -0.49), (-0.71, 1.20, -0.49), (-0.14, 1.20, -0.67), (-0.14, 0.00, -0.67)],
        [(-0.14, 0.00, -0.67), (-0.14, 1.20, -0.67), (0.04, 1.20, -0.09), (0.04, 0.00, -0.09)],
        [(0.04, 0.00, -0.09), (0.04, 1.20, -0.09), (-0.53, 1.20, 0.09), (-0.53, 0.00, 0.09)],
        [(-0.53, 0.00, 0.09), (0.04, 0.00, -0.09), (-0.14, 0.00, -0.67), (-0.71, 0.00, -0.49)]],
    "light": [[(-0.24, 1.98, 0.16), (-0.24, 1.98, -0.22), (0.23, 1.98, -0.22), (0.23, 1.98, 0.16)]],
}
CBOX_GROUPS = ["floor", "ceiling", "backWall", "rightWall", "leftWall", "shortBox", "tallBox", "light"]


def c2_cornell():
    s = SceneDesc()
    for g in CBOX_GROUPS:
        quads = _CBOX[g]
        verts = np.array([v for q in quads for v in q], dtype=np.float32)
        tris = np.array([t for k in range(len(quads)) for t in ((4 * k, 4 * k + 1, 4 * k + 2), (4 * k, 4 * k + 2, 4 * k + 3))], dtype=np.uint32)
        s.add_instance(s.add_mesh(verts, tris), mask=255, opaque=True)
    return s


def c2_primary_rays(w=1024, h=1024, seed=0xC0FFEE):
    """Camera of path_tracer.rs:291-306 with the per-pixel LCG jitter of :271-279 (seed image = splitmix64(seed + pixel))."""
    n = w * h
    idx = np.arange(n, dtype=np.uint64)
    z = idx + np.uint64(seed) + np.uint64(0x9E3779B97F4A7C15)
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    state = ((z ^ (z >> np.uint64(31))) & np.uint64(0xFFFFFFFF)).astype(np.uint32)

    def lcg(st):
        st = (np.uint32(1664525) * st + np.uint32(1013904223)).astype(np.uint32)
        return st, (st & np.uint32(0x00FFFFFF)).astype(np.float32) * np.float32(1.0 / 16777216.0)
    state, rx = lcg(state)
    state, ry = lcg(state)
    x = (np.arange(n) % w).astype(np.float32)
    y = (np.arange(n) // w).astype(np.float32)
    frame = np.float32(min(w, h))
    px = (x + rx) / frame * np.float32(2.0) - np.float32(1.0)
    py = -((y + ry) / frame * np.float32(2.0) - np.float32(1.0))
    t = np.float32(np.tan(0.5 * 27.8 * np.pi / 180.0))
    origin = np.array([-0.01, 0.995, 5.0], np.float32)
    pixel = origin + np.stack([px * t, py * t, np.full(n, -1.0, np.float32)], 1)
    d = pixel - origin
    d = (d / np.sqrt((d * d).sum(1, keepdims=True, dtype=np.float32))).astype(np.float32)
    return make_rays(np.broadcast_to(origin, (n, 3)), d, np.float32(0.0), np.float32(3.4028234663852886e38))


# ---- C3: random triangle soup + incoherent rays -----------------------------------------------------
def random_soup(n_tris, seed=0x5EED0001, extent=0.01, indexed=False):
    """centroid ~ U[0,1]^3, two edge vectors ~ U[-extent,extent]^3, unindexed (idx = 3i,3i+1,3i+2)."""
    rng = np.random.default_rng(seed)
    c = rng.random((n_tris, 3), dtype=np.float32)
    e1 = (rng.random((n_tris, 3), dtype=np.float32) * 2 - 1) * np.float32(extent)
    e2 = (rng.random((n_tris, 3), dtype=np.float32) * 2 - 1) * np.float32(extent)
    v = np.empty((n_tris, 3, 3), np.float32)
    v[:, 0] = c - (e1 + e2) / np.float32(3)
    v[:, 1] = v[:, 0] + e1
    v[:, 2] = v[:, 0] + e2
    verts = v.reshape(-1, 3)
    tris = np.arange(3 * n_tris, dtype=np.uint32).reshape(-1, 3)
    return verts, tris


def c3_soup(n_tris=1_000_000, seed=0x5EED0001):
    s = SceneDesc()
    s.add_instance(s.add_mesh(*random_soup(n_tris, seed)))
    return s


def incoherent_rays(n, seed=0x5EED0002, tmin=1e-4, tmax=1e30, lo=0.0, hi=1.0):
    """origin ~ U[lo,hi]^3, direction uniform on the sphere."""
    rng = np.random.default_rng(seed)
    o = (rng.random((n, 3), dtype=np.float32) * np.float32(hi - lo) + np.float32(lo)).astype(np.float32)
    z = rng.random(n, dtype=np.float32) * 2 - 1
    phi = rng.random(n, dtype=np.float32) * np.float32(2 * np.pi)
    r = np.sqrt(np.maximum(0, 1 - z * z)).astype(np.float32)
    d = np.stack([r * np.cos(phi), r * np.sin(phi), z], 1).astype(np.float32)
    return make_rays(o, d, np.float32(tmin), np.float32(tmax))


def shadow_rays_from_hits(rays, hits, seed=0x5EED0004):
    """From each hit point toward a U[0,1]^3 target, tmax = dist*(1-1e-4) (C3's trace_any set); misses aim from the ray origin."""
    rng = np.random.default_rng(seed)
    n = rays.shape[0]
    valid = hits["inst"] != 0xFFFFFFFF
    t = np.where(valid, hits["committed_ray_t"], np.float32(0)).astype(np.float32)
    p = rays["orig"] + rays["dir"] * t[:, None]
    target = rng.random((n, 3), dtype=np.float32)
    d = target - p
    dist = np.sqrt((d * d).sum(1, dtype=np.float32)).astype(np.float32)
    dist = np.maximum(dist, np.float32(1e-6))
    d = (d / dist[:, None]).astype(np.float32)
    return make_rays(p.astype(np.float32), d, np.float32(1e-4), (dist * np.float32(1 - 1e-4)).astype(np.float32))


# ---- C4-style terrain ---------------------------------------------------------------------------
def terrain(nx, seed=0x5EED0003, frame=0):
    """nx x nx vertex height field over [0,1]^2 (2(nx-1)^2 triangles), h = 0.1 * 5-octave value noise (+ per-frame ripple)."""
    rng = np.random.default_rng(seed)
    u = np.linspace(0, 1, nx, dtype=np.float32)
    x, z = np.meshgrid(u, u)
    h = np.zeros((nx, nx), np.float32)
    amp, freq = 0.5, 4
    for _ in range(5):
        g = rng.random((freq + 2, freq + 2), dtype=np.float32)
        fx, fz = x * freq, z * freq
        ix, iz = np.minimum(fx.astype(np.int32), freq), np.minimum(fz.astype(np.int32), freq)
        tx, tz = fx - ix, fz - iz
        tx, tz = tx * tx * (3 - 2 * tx), tz * tz * (3 - 2 * tz)
        a = g[iz, ix] * (1 - tx) + g[iz, ix + 1] * tx
        b = g[iz + 1, ix] * (1 - tx) + g[iz + 1, ix + 1] * tx
        h += np.float32(amp) * (a * (1 - tz) + b * tz)
        amp *= 0.5
        freq *= 2
    y = np.float32(0.1) * h
    if frame:
        y = y + np.float32(0.01) * np.sin(np.float32(40.0) * x + np.float32(0.3 * frame)).astype(np.float32)
    verts = np.stack([x, y.astype(np.float32), z], -1).reshape(-1, 3).astype(np.float32)
    i = np.arange(nx - 1, dtype=np.uint32)
    a = (i[:, None] * nx + i[None, :]).reshape(-1)
    tris = np.concatenate([np.stack([a, a + nx, a + 1], 1), np.stack([a + 1, a + nx, a + nx + 1], 1)], 0).astype(np.uint32)
    return verts, tris


def rotation_y(deg):
    c, s = np.cos(np.radians(deg)), np.sin(np.radians(deg))
    return np.array([[c, 0, s, 0], [0, 1, 0, 0], [-s, 0, c, 0]], dtype=np.float32)


def instanced_scene(n_tris_per_mesh=2000, n_instances=10, seed=7):
    """C5-shaped: one mesh instanced on a 5x2 grid with yaw 36*k degrees, plus a second small mesh, varied masks."""
    s = SceneDesc()
    m0 = s.add_mesh(*random_soup(n_tris_per_mesh, seed, extent=0.05))
    m1 = s.add_mesh(*random_soup(max(4, n_tris_per_mesh // 10), seed + 1, extent=0.08))
    for k in range(n_instances):
        t = rotation_y(36.0 * k)
        t[:, 3] = [1.5 * (k % 5), 0.25 * (k % 3), 1.5 * (k // 5)]
        s.add_instance(m0 if k % 3 else m1, t, mask=0xFF if k % 4 else 0x0F, user_id=100 + k)
    return s
