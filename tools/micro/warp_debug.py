import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import luisa_compute_rs_b200 as lc
from luisa_compute_rs_b200 import ir
from luisa_compute_rs_b200.ir import Func
dev = lc.Context().create_device("b200")
rng = np.random.default_rng(42)
n = 128 * 50
x = rng.integers(0, 2**20, n, dtype=np.uint32)
k = ir.KernelBuilder(block_size=(128, 1, 1))
a, wout, facc = k.arg_buffer(k.u32), k.arg_buffer(k.u32), k.arg_buffer(k.f32)
MODE = int(sys.argv[1]) if len(sys.argv) > 1 else 0
def body():
    i = k.dispatch_id().x
    xv = a.read(i)
    if MODE & 1:
        k.call(Func.AtomicFetchMax, [facc, k.u(1), xv.cast(k.f32)], k.f32)
    if MODE & 2:
        k.call(Func.AtomicFetchAdd, [facc, k.u(0), (xv & k.u(255)).cast(k.f32)], k.f32)
    if MODE & 4:
        k.if_(xv.eq(12345), lambda: k.call(Func.AtomicFetchAdd, [facc, k.u(0), k.f(1.0)], k.f32))
    vals = [k.call(Func.WarpActiveSum, [xv & k.u(1023)], k.u32), k.call(Func.WarpPrefixSum, [xv & k.u(1023)], k.u32), k.call(Func.WarpActiveMax, [xv], k.u32),
            k.call(Func.WarpActiveCountBits, [(xv & k.u(1)).eq(1)], k.u32), k.call(Func.WarpReadFirstLane, [xv], k.u32)]
    lane = k.call(Func.WarpLaneId, [], k.u32)
    vals.append(k.call(Func.WarpReadLaneAt, [xv, (lane + k.u(5)) & k.u(31)], k.u32))
    vals.append(lane)
    for j, v in enumerate(vals):
        wout.write(i * k.u(8) + k.u(j), v)
k.body(body); k.finish()
sh = dev.create_shader(C.addressof(k.km), keep=k)
ab = dev.create_buffer_from_array(x); wb = dev.create_buffer(n * 8, 4)
fb = dev.create_buffer_from_array(np.array([0, -1, 1e30], np.float32))
sh.dispatch((n,), ab, wb, fb)
print("MODE", MODE)
if MODE & 8:
    print(C.string_at(dev.lib.lc_b200_ir_lower_source(C.addressof(k.km))).decode()[-1800:])
got = wb.view().to_numpy(np.uint32).reshape(-1, 32, 8)
w = x.reshape(-1, 32).astype(np.uint64); low = w & 1023
want = [np.broadcast_to(low.sum(1, keepdims=True), w.shape), np.cumsum(low, 1) - low, np.broadcast_to(w.max(1, keepdims=True), w.shape),
        np.broadcast_to((w & 1).sum(1, keepdims=True), w.shape), np.broadcast_to(w[:, :1], w.shape), np.roll(w, -5, axis=1), np.broadcast_to(np.arange(32), w.shape)]
for j, (name, wv) in enumerate(zip(["sum", "prefix", "max", "count", "first", "lane_at", "lane"], want)):
    bad = got[:, :, j] != wv
    print(name, "mismatches", int(bad.sum()), "example got/want", got[1, :4, j], np.asarray(wv)[1, :4])
dev.close()
