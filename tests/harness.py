"""Builds the same SceneDesc on the b200 device (through the C ABI) and in the oracle, and compares hit buffers."""
import numpy as np

import luisa_compute_rs_b200 as lc
import oracle_lib as ol


class DeviceScene:
    def __init__(self, device, desc, option=None, vertex_stride=12):
        self.device = device
        self.buffers = []
        self.meshes = []
        for verts, tris in desc.meshes:
            if vertex_stride == 12:
                vb = device.create_buffer_from_array(verts)
            else:  # Float3-style 16-byte stride (ir.rs:234-263)
                padded = np.zeros((verts.shape[0], vertex_stride // 4), np.float32)
                padded[:, :3] = verts
                vb = device.create_buffer_from_array(padded)
            ib = device.create_buffer_from_array(tris if tris.shape[0] else np.zeros((0, 3), np.uint32))
            self.buffers += [vb, ib]
            m = device.create_mesh(vb.view(), ib.view(), option or lc.AccelOption())
            m.build(lc.AccelBuildRequest.FORCE_BUILD)
            self.meshes.append(m)
        self.accel = device.create_accel(option or lc.AccelOption())
        for inst in desc.instances:
            t = np.eye(4, dtype=np.float32)
            t[:3, :] = inst["transform"]
            self.accel.push_mesh(self.meshes[inst["mesh"]], t, inst["mask"], inst["opaque"])
            if inst["user_id"]:
                self.accel.set_user_id_on_update(len(self.accel.instance_handles) - 1, inst["user_id"])
        self.accel.build(lc.AccelBuildRequest.FORCE_BUILD)

    def trace_closest(self, rays, mask=0xFF):
        """device-buffer path: upload rays, lc_b200_trace_closest, download hits"""
        n = rays.shape[0]
        rb = self.device.create_buffer(max(n, 1), 32, 16)
        hb = self.device.create_buffer(max(n, 1), 24, 8)
        hits = np.zeros(n, dtype=lc.SurfaceHit)
        if n:
            rb.view(0, n).copy_from(rays)
            self.accel.intersect(rb.view(0, n), hb.view(0, n), n, mask)
            hb.view(0, n).copy_to(hits)
        rb.destroy(); hb.destroy()
        return hits

    def trace_any(self, rays, mask=0xFF):
        n = rays.shape[0]
        rb = self.device.create_buffer(max(n, 1), 32, 16)
        ob = self.device.create_buffer(max(n, 1), 4, 4)
        occ = np.zeros(n, dtype=np.uint32)
        if n:
            rb.view(0, n).copy_from(rays)
            self.accel.intersect_any(rb.view(0, n), ob.view(0, n), n, mask)
            ob.view(0, n).copy_to(occ)
        rb.destroy(); ob.destroy()
        return occ

    def trace_dsl(self, rays, any_hit=False, mask=0xFF, block_size=(64, 1, 1)):
        """the reference call path: a DSL kernel `hits.write(i, accel.intersect(rays.read(i), mask))` as an ir::KernelModule through
        create_shader + ShaderDispatch (examples_ir.trace_buffer_kernel)"""
        import ctypes as C
        from luisa_compute_rs_b200 import examples_ir
        n = rays.shape[0]
        rb = self.device.create_buffer(max(n, 1), 32, 16)
        ob = self.device.create_buffer(max(n, 1), 4, 4) if any_hit else self.device.create_buffer(max(n, 1), 24, 8)
        out = np.zeros(n, dtype=np.uint32 if any_hit else lc.SurfaceHit)
        k = examples_ir.trace_buffer_kernel(any_hit=any_hit, mask=mask, block_size=block_size)
        sh = self.device.create_shader(C.addressof(k.km), keep=k)
        if n:
            rb.view(0, n).copy_from(rays)
            sh.dispatch((n, 1, 1), rb, ob, self.accel)
            ob.view(0, n).copy_to(out)
        sh.destroy(); rb.destroy(); ob.destroy()
        return out

    def destroy(self):
        self.accel.destroy()
        for m in self.meshes:
            m.destroy()
        for b in self.buffers:
            b.destroy()


def assert_hits_equal(got, want, what=""):
    """Bit-exact on every field of SurfaceHit (inst, prim, bary, t): the GPU implements the oracle's canonical arithmetic."""
    assert got.shape == want.shape
    for f in ("inst", "prim"):
        bad = np.nonzero(got[f] != want[f])[0]
        assert bad.size == 0, f"{what}: {bad.size} of {got.shape[0]} rays differ in {f}; first {bad[:5]}: got {got[f][bad[:5]]} want {want[f][bad[:5]]}"
    assert np.array_equal(got["committed_ray_t"].view(np.uint32), want["committed_ray_t"].view(np.uint32)), f"{what}: t differs bitwise"
    assert np.array_equal(got["bary"].view(np.uint32), want["bary"].view(np.uint32)), f"{what}: barycentrics differ bitwise"


def compare_with_truth(hits, truth, ambiguous, rel_tol=1e-5):
    """north_star tolerance: inst/prim identical wherever the f64 answer is unambiguous; t and bary within 1e-5 relative.
    Returns (tie_rate, mismatches_outside_ties)."""
    clear = ambiguous == 0
    same = (hits["inst"] == truth["inst"]) & (hits["prim"] == truth["prim"])
    bad = np.nonzero(clear & ~same)[0]
    v = clear & same & (truth["inst"] != 0xFFFFFFFF)
    t_err = np.abs(hits["committed_ray_t"][v].astype(np.float64) - truth["committed_ray_t"][v]) / np.maximum(np.abs(truth["committed_ray_t"][v]), 1e-30)
    b_err = np.abs(hits["bary"][v].astype(np.float64) - truth["bary"][v]).max(initial=0.0)
    return dict(tie_rate=float(ambiguous.mean()) if ambiguous.size else 0.0, mismatches=int(bad.size), disagree_in_ties=int((~clear & ~same).sum()),
                t_rel_err=float(t_err.max(initial=0.0)), bary_abs_err=float(b_err))
