"""Host-side mirror of luisa_compute::rtx (luisa_compute/src/rtx.rs) for the B200 device.

`Mesh` / `Accel` record exactly the `api::MeshBuildCommand` / `api::AccelBuildCommand` +
`AccelBuildModification`s the Rust frontend records (rtx.rs:127-147, 154-311).  The device-side
DSL builtins `AccelVar::intersect` / `intersect_any` (rtx.rs:774-817; deprecated aliases
`trace_closest` / `trace_any`, rtx.rs:819-870) are exposed in batch form over buffers of rays.
"""
import ctypes as C

import numpy as np

from . import _abi as abi
from ._abi import AccelOption  # noqa: F401  (re-export, rtx.rs re-exports api::AccelOption)
from .runtime import HostCommand, LuisaError


class AccelBuildRequest:  # api_types lib.rs:188-193
    PREFER_UPDATE = 0
    FORCE_BUILD = 1


class AccelUsageHint:  # api_types lib.rs:196-202
    FAST_TRACE = 0
    FAST_BUILD = 1


# rtx.rs:329-338 — 32 bytes, align 16
Ray = np.dtype([("orig", "<f4", (3,)), ("tmin", "<f4"), ("dir", "<f4", (3,)), ("tmax", "<f4")])
# rtx.rs:356-366 — 24 bytes, align 8
SurfaceHit = np.dtype([("inst", "<u4"), ("prim", "<u4"), ("bary", "<f4", (2,)), ("committed_ray_t", "<f4"), ("_pad", "<u4")])
# rtx.rs:536
Index = np.dtype(("<u4", (3,)))
# rtx.rs:477-485 — 24 bytes, align 8; hit_type: 0 miss, 1 surface (triangle), 2 procedural (rtx.rs:486-493)
CommittedHit = np.dtype([("inst", "<u4"), ("prim", "<u4"), ("bary", "<f4", (2,)), ("hit_type", "<u4"), ("committed_ray_t", "<f4")])
assert Ray.itemsize == 32 and SurfaceHit.itemsize == 24 and CommittedHit.itemsize == 24


class HitType:  # rtx.rs:486-493
    MISS = 0
    SURFACE = 1
    PROCEDURAL = 2


class SurfaceCandidateFilter:
    """Stand-in for the `on_surface_hit(|candidate| ...)` closure of `AccelVar::traverse` (rtx.rs:871-902) in the batch
    form: a device-side pure predicate over the candidate (inst, prim, bary) that decides `candidate.commit()`."""

    def __init__(self, kind=abi.FILTER_COMMIT_ALL, radius=0.0, bits=None, first_bit=None):
        self.kind, self.radius, self.bits, self.first_bit = kind, radius, bits, first_bit

    @staticmethod
    def commit_all():
        return SurfaceCandidateFilter(abi.FILTER_COMMIT_ALL)

    @staticmethod
    def reject_all():
        return SurfaceCandidateFilter(abi.FILTER_REJECT_ALL)

    @staticmethod
    def bary_disc(radius):
        """examples/ray_query.rs:148-162"""
        return SurfaceCandidateFilter(abi.FILTER_BARY_DISC, radius=radius)

    @staticmethod
    def prim_bits(bits_buffer, first_bit_buffer):
        return SurfaceCandidateFilter(abi.FILTER_PRIM_BITS, bits=bits_buffer, first_bit=first_bit_buffer)

    def _c(self):
        f = abi.CandidateFilter(self.kind, self.radius, abi.Handle(0), abi.Handle(0))
        if self.kind == abi.FILTER_PRIM_BITS:
            f.bits, f.first_bit = self.bits.handle, self.first_bit.handle
        return f

INVALID = 0xFFFFFFFF


def make_rays(orig, direction, tmin, tmax):
    """`Ray::new_expr(o, tmin, d, tmax)` for arrays."""
    orig = np.asarray(orig, dtype=np.float32)
    n = orig.shape[0]
    r = np.empty(n, dtype=Ray)
    r["orig"] = orig
    r["dir"] = np.asarray(direction, dtype=np.float32)
    r["tmin"] = tmin
    r["tmax"] = tmax
    return r


def hit_valid(hits):
    """`SurfaceHit::valid` (rtx.rs:573-578): inst != u32::MAX."""
    return hits["inst"] != INVALID


def affine_from_mat4(m):
    """`Mat4::into_affine3x4` (lang/types/vector.rs:645-660): rows 0..2 of the 4x4, row-major, 12 floats."""
    m = np.asarray(m, dtype=np.float32)
    if m.shape == (4, 4):
        return np.ascontiguousarray(m[:3, :]).reshape(12)
    if m.shape == (3, 4):
        return np.ascontiguousarray(m).reshape(12)
    if m.shape == (12,):
        return np.ascontiguousarray(m)
    raise LuisaError("transform must be 4x4, 3x4 or 12 floats")


class Mesh:
    """`Device::create_mesh(vertex_view, index_view, option)` (runtime.rs:662-689)."""

    def __init__(self, device, vertex_view, index_view, option=None):
        self.device = device
        self.option = option or AccelOption()
        self.vertex_view = vertex_view
        self.index_view = index_view
        self.vertex_stride = vertex_view.buffer.stride
        self.index_stride = index_view.buffer.stride
        info = device.iface.create_mesh(device.handle, C.byref(self.option))
        self.handle = abi.Handle(info.handle)
        self._alive = True

    def build_async(self, request=AccelBuildRequest.FORCE_BUILD):
        cmd = abi.Command()
        cmd.tag = abi.CMD_MESH_BUILD
        cmd.u.mesh_build = abi.CmdMeshBuild(
            self.handle, request, self.vertex_view.buffer.handle, self.vertex_view.offset, self.vertex_view.size, self.vertex_stride,
            self.index_view.buffer.handle, self.index_view.offset, self.index_view.size, self.index_stride)
        return HostCommand(cmd, keep=[self, self.vertex_view.buffer, self.index_view.buffer])

    def build(self, request=AccelBuildRequest.FORCE_BUILD):
        s = self.device.default_stream()
        s.submit([self.build_async(request)])
        s.synchronize()

    def stats(self):
        st = abi.BuildStats()
        self.device.lib.lc_b200_mesh_stats(self.device.handle, self.handle, C.byref(st))
        return st.as_dict()

    def destroy(self):
        if self._alive and not self.device._closed:
            self.device.iface.destroy_mesh(self.device.handle, self.handle)
        self._alive = False


class ProceduralPrimitive:
    """`Device::create_procedural_primitive(aabb_view, option)` (runtime.rs:702-720; rtx.rs:65-95): a BLAS over user AABBs
    ({min: [f32; 3], max: [f32; 3]}, 24 B).  Hits come from the `on_procedural_hit` block of a RayQuery."""

    def __init__(self, device, aabb_view, option=None):
        self.device = device
        self.option = option or AccelOption()
        self.aabb_view = aabb_view
        if aabb_view.buffer.stride != 24:
            raise LuisaError("procedural primitives need a buffer of Aabb {min, max} records (24 bytes)")
        info = device.iface.create_procedural_primitive(device.handle, C.byref(self.option))
        self.handle = abi.Handle(info.handle)
        self._alive = True

    def build_async(self, request=AccelBuildRequest.FORCE_BUILD):
        cmd = abi.Command()
        cmd.tag = abi.CMD_PROCEDURAL_BUILD
        cmd.u.procedural_build = abi.CmdProceduralBuild(self.handle, request, self.aabb_view.buffer.handle, self.aabb_view.offset, self.aabb_view.size // 24)
        return HostCommand(cmd, keep=[self, self.aabb_view.buffer])

    def build(self, request=AccelBuildRequest.FORCE_BUILD):
        s = self.device.default_stream()
        s.submit([self.build_async(request)])
        s.synchronize()

    def destroy(self):
        if self._alive and not self.device._closed:
            self.device.iface.destroy_procedural_primitive(self.device.handle, self.handle)
        self._alive = False


class CurveBasis:
    """api::CurveBasis (api_types:196-202)"""
    PIECEWISE_LINEAR, CUBIC_BSPLINE, CATMULL_ROM, BEZIER = 0, 1, 2, 3


class Curve:
    """A curve geometry behind `create_curve` / `CurveBuildCommand` (api_types:618-631; backend: GeometryImpl::build_curve,
    cpu/accel.rs:142-203).  The Rust frontend of the reference has the device-side half only (SurfaceHit::is_curve / curve_parameter,
    rtx/curve.rs evaluators); the host type follows the C++ runtime's `Device::create_curve(basis, control_points, segments)`.
    control points: a buffer of float4 {x, y, z, radius} (stride >= 16); segments: a buffer of u32 first-control-point indices."""

    def __init__(self, device, basis, cp_view, seg_view, option=None):
        self.device = device
        self.option = option or AccelOption()
        self.basis = int(basis)
        self.cp_view, self.seg_view = cp_view, seg_view
        if cp_view.buffer.stride < 16:
            raise LuisaError("cp buffer stride must be >= 16")
        if seg_view.buffer.stride != 4:
            raise LuisaError("curve segments are u32 control point indices")
        info = device.iface.create_curve(device.handle, C.byref(self.option))
        self.handle = abi.Handle(info.handle)
        self._alive = True

    def build_async(self, request=AccelBuildRequest.FORCE_BUILD):
        cmd = abi.Command()
        cmd.tag = abi.CMD_CURVE_BUILD
        cp, sg = self.cp_view, self.seg_view
        cmd.u.curve_build = abi.CmdCurveBuild(self.handle, request, self.basis, cp.size // cp.buffer.stride, sg.size // 4,
                                              cp.buffer.handle, cp.offset, cp.buffer.stride, sg.buffer.handle, sg.offset)
        return HostCommand(cmd, keep=[self, cp.buffer, sg.buffer])

    def build(self, request=AccelBuildRequest.FORCE_BUILD):
        s = self.device.default_stream()
        s.submit([self.build_async(request)])
        s.synchronize()

    def destroy(self):
        if self._alive and not self.device._closed:
            self.device.iface.destroy_curve(self.device.handle, self.handle)
        self._alive = False


class Accel:
    """`Device::create_accel(option)` (runtime.rs:690-701) and `rtx::Accel` (rtx.rs:154-311)."""

    def __init__(self, device, option=None):
        self.device = device
        self.option = option or AccelOption()
        info = device.iface.create_accel(device.handle, C.byref(self.option))
        self.handle = abi.Handle(info.handle)
        self.instance_handles = []
        self.modifications = {}
        self._alive = True

    # rtx.rs:154-189.  NB the reference sets `index = modifications.len()`, which is only right while the
    # pending map holds one entry per instance; we record the slot index the call means (instance_handles.len()).
    def _push_handle(self, mesh, transform, ray_mask, opaque, user_id=0):
        flags = abi.MOD_PRIMITIVE | abi.MOD_TRANSFORM | abi.MOD_VISIBILITY | abi.MOD_USER_ID
        flags |= abi.MOD_OPAQUE_ON if opaque else abi.MOD_OPAQUE_OFF
        index = len(self.instance_handles)
        m = abi.AccelModification(index, user_id, flags, ray_mask, mesh.handle.id, (C.c_float * 12)(*affine_from_mat4(transform)))
        self.modifications[index] = m
        self.instance_handles.append(mesh)

    # rtx.rs:190-223: PRIMITIVE without TRANSFORM (the backend applies the affine of a PRIMITIVE modification anyway)
    def _set_handle(self, index, mesh, transform, ray_mask, opaque, user_id=0):
        flags = abi.MOD_PRIMITIVE | abi.MOD_VISIBILITY | abi.MOD_USER_ID
        flags |= abi.MOD_OPAQUE_ON if opaque else abi.MOD_OPAQUE_OFF
        m = abi.AccelModification(index, user_id, flags, ray_mask, mesh.handle.id, (C.c_float * 12)(*affine_from_mat4(transform)))
        self.modifications[index] = m
        self.instance_handles[index] = mesh

    def push_mesh(self, mesh, transform=None, ray_mask=0xFF, opaque=True):
        self._push_handle(mesh, np.eye(4, dtype=np.float32) if transform is None else transform, ray_mask, opaque)

    def push_procedural_primitive(self, prim, transform=None, ray_mask=0xFF):
        """rtx.rs:236-247: procedural instances are never opaque (every candidate goes through the callback)."""
        self._push_handle(prim, np.eye(4, dtype=np.float32) if transform is None else transform, ray_mask, False)

    def push_curve(self, curve, transform=None, ray_mask=0xFF, opaque=True):
        self._push_handle(curve, np.eye(4, dtype=np.float32) if transform is None else transform, ray_mask, opaque)

    def set_mesh(self, index, mesh, transform=None, ray_mask=0xFF, opaque=True):
        self._set_handle(index, mesh, np.eye(4, dtype=np.float32) if transform is None else transform, ray_mask, opaque)

    def pop(self):
        n = len(self.instance_handles)
        self.modifications.pop(n, None)
        self.instance_handles.pop()

    def _modify(self, index, flag, **kw):
        m = self.modifications.get(index)
        if m is None:
            m = abi.AccelModification(index, 0, 0, 0, 0, (C.c_float * 12)(1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0))
            self.modifications[index] = m
        m.flags |= flag
        for k, v in kw.items():
            setattr(m, k, v)

    def set_transform_on_update(self, index, transform):
        self._modify(index, abi.MOD_TRANSFORM, affine=(C.c_float * 12)(*affine_from_mat4(transform)))

    def set_visibility_on_update(self, index, mask):
        self._modify(index, abi.MOD_VISIBILITY, visibility=mask)

    def set_user_id_on_update(self, index, user_id):
        self._modify(index, abi.MOD_USER_ID, user_id=user_id)

    def build_async(self, request=AccelBuildRequest.FORCE_BUILD, instance_buffer_only=False):
        """`Accel::build_async` (rtx.rs:287-311; the Rust frontend always passes update_instance_buffer_only = false, the C++ one
        exposes it as Accel::update_instance_buffer, runtime/rtx/accel.cpp:47-63): apply the pending modifications to the instance
        table and — unless instance_buffer_only — rebuild the TLAS."""
        mods = list(self.modifications.values())  # drained, like HashMap::drain (rtx.rs:295)
        self.modifications = {}
        arr = (abi.AccelModification * max(len(mods), 1))(*mods)
        cmd = abi.Command()
        cmd.tag = abi.CMD_ACCEL_BUILD
        cmd.u.accel_build = abi.CmdAccelBuild(self.handle, request, len(self.instance_handles), arr, len(mods), bool(instance_buffer_only))
        return HostCommand(cmd, keep=[self, arr] + list(self.instance_handles))

    def build(self, request=AccelBuildRequest.FORCE_BUILD, instance_buffer_only=False):
        s = self.device.default_stream()
        s.submit([self.build_async(request, instance_buffer_only)])
        s.synchronize()

    # ---- ray queries: batch form of AccelVar::intersect / intersect_any --------------------------
    def intersect(self, rays, hits, count=None, mask=0xFF, stream=None):
        """Closest hit for `count` rays of buffer `rays` (Ray, 32 B) into buffer `hits` (SurfaceHit, 24 B); asynchronous."""
        s = stream or self.device.default_stream()
        rv, hv = _as_view(rays), _as_view(hits)
        n = rv.size // 32 if count is None else count
        self.device.lib.lc_b200_trace_closest(self.device.handle, s.handle, self.handle, rv.buffer.handle, rv.offset, hv.buffer.handle, hv.offset, n, mask)

    def intersect_any(self, rays, occluded, count=None, mask=0xFF, stream=None):
        """Any-hit for `count` rays into a buffer of uint32 (1 = occluded); asynchronous."""
        s = stream or self.device.default_stream()
        rv, ov = _as_view(rays), _as_view(occluded)
        n = rv.size // 32 if count is None else count
        self.device.lib.lc_b200_trace_any(self.device.handle, s.handle, self.handle, rv.buffer.handle, rv.offset, ov.buffer.handle, ov.offset, n, mask)

    def traverse(self, rays, committed, count=None, mask=0xFF, on_surface_hit=None, stream=None, terminate_on_first=False):
        """Batch form of `accel.traverse(ray, opts).on_surface_hit(f).trace()` (rtx.rs:871-902, 678-755): CommittedHit per ray."""
        s = stream or self.device.default_stream()
        rv, hv = _as_view(rays), _as_view(committed)
        n = rv.size // 32 if count is None else count
        f = (on_surface_hit or SurfaceCandidateFilter.commit_all())._c()
        self.device.lib.lc_b200_ray_query(self.device.handle, s.handle, self.handle, rv.buffer.handle, rv.offset, hv.buffer.handle, hv.offset, n, mask,
                                          terminate_on_first, C.byref(f))

    def traverse_any(self, rays, committed, count=None, mask=0xFF, on_surface_hit=None, stream=None):
        """`accel.traverse_any` (rtx.rs:884-902): terminate on the first committed hit."""
        self.traverse(rays, committed, count, mask, on_surface_hit, stream, terminate_on_first=True)

    trace_closest = intersect   # rtx.rs:819-843 deprecated alias
    trace_any = intersect_any   # rtx.rs:844-870 deprecated alias

    def intersect_counted(self, rays, hits, count=None, mask=0xFF, stream=None):
        """Same as intersect() through the instrumented kernel; returns node / triangle visit sums (synchronous)."""
        s = stream or self.device.default_stream()
        rv, hv = _as_view(rays), _as_view(hits)
        n = rv.size // 32 if count is None else count
        ctr = abi.TraceCounters()
        self.device.lib.lc_b200_trace_closest_counted(self.device.handle, s.handle, self.handle, rv.buffer.handle, rv.offset, hv.buffer.handle, hv.offset, n, mask, C.byref(ctr))
        return {k: getattr(ctr, k) for k, _ in ctr._fields_}

    def intersect_host(self, rays, mask=0xFF, out=None):
        """Host arrays in, host arrays out (H2D, trace, D2H, synchronous): the end-to-end call."""
        rays = _check_rays(rays)
        n = rays.shape[0]
        hits = np.empty(n, dtype=SurfaceHit) if out is None else out
        self.device.lib.lc_b200_trace_closest_host(self.device.handle, self.handle, rays.ctypes.data, hits.ctypes.data, n, mask)
        return hits

    def intersect_any_host(self, rays, mask=0xFF, out=None):
        rays = _check_rays(rays)
        n = rays.shape[0]
        occ = np.empty(n, dtype=np.uint32) if out is None else out
        self.device.lib.lc_b200_trace_any_host(self.device.handle, self.handle, rays.ctypes.data, occ.ctypes.data, n, mask)
        return occ

    def intersect_host_ptr(self, rays_ptr, hits_ptr, n, mask=0xFF):
        """Raw-pointer form of intersect_host (pinned torch tensors: pass tensor.data_ptr())."""
        self.device.lib.lc_b200_trace_closest_host(self.device.handle, self.handle, rays_ptr, hits_ptr, n, mask)

    def intersect_any_host_ptr(self, rays_ptr, occ_ptr, n, mask=0xFF):
        self.device.lib.lc_b200_trace_any_host(self.device.handle, self.handle, rays_ptr, occ_ptr, n, mask)

    # ---- instance accessors (AccelVar::instance_transform etc., rtx.rs:918-1000; cpu/accel.rs:537-558) ----
    def instance_transform(self, index):
        out = (C.c_float * 12)()
        self.device.lib.lc_b200_instance_transform(self.device.handle, self.handle, index, out)
        return np.array(out, dtype=np.float32).reshape(3, 4)

    def instance_user_id(self, index):
        return self.device.lib.lc_b200_instance_user_id(self.device.handle, self.handle, index)

    def instance_visibility_mask(self, index):
        return self.device.lib.lc_b200_instance_visibility_mask(self.device.handle, self.handle, index)

    def stats(self):
        st = abi.BuildStats()
        self.device.lib.lc_b200_accel_stats(self.device.handle, self.handle, C.byref(st))
        return st.as_dict()

    def destroy(self):
        if self._alive and not self.device._closed:
            self.device.iface.destroy_accel(self.device.handle, self.handle)
        self._alive = False


def _as_view(b):
    return b.view() if hasattr(b, "count") else b


def _check_rays(rays):
    rays = np.ascontiguousarray(rays)
    if rays.dtype != Ray:
        raise LuisaError("rays must have dtype rtx.Ray")
    return rays


def offset_ray_origin(p, n):
    """`offset_ray_origin` (rtx.rs:517-535), the self-intersection-avoiding origin offset, for float32 arrays (N,3)."""
    p = np.asarray(p, dtype=np.float32)
    n = np.asarray(n, dtype=np.float32)
    origin, float_scale, int_scale = np.float32(1.0 / 32.0), np.float32(1.0 / 65536.0), np.float32(256.0)
    of_i = (int_scale * n).astype(np.int32)
    p_i = p.view(np.int32) + np.where(p < 0, -of_i, of_i)
    return np.where(np.abs(p) < origin, p + float_scale * n, p_i.view(np.float32)).astype(np.float32)
