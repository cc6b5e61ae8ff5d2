#!/usr/bin/env python
"""Where the time of the host-buffer pipeline (bench.py e2e: BufferUpload | ShaderDispatch | BufferDownload on three streams) goes:
per-chunk CUDA-event timestamps on each of the three streams + host-side time spent inside each submit.  Development aid."""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import luisa_compute_rs_b200 as lc  # noqa: E402
import scenes  # noqa: E402
from luisa_compute_rs_b200 import examples_ir  # noqa: E402


def main():
    chunk = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 21
    n = 1 << 24
    dev = lc.Context().create_device("b200")
    verts, tris = scenes.random_soup(1_000_000, 0x5EED0001)
    vb, ib = dev.create_buffer_from_array(verts), dev.create_buffer_from_array(tris)
    mesh = dev.create_mesh(vb.view(), ib.view(), lc.AccelOption()); mesh.build()
    accel = dev.create_accel(); accel.push_mesh(mesh); accel.build()
    k = examples_ir.trace_buffer_kernel()
    shader = dev.create_shader(C.addressof(k.km), keep=k)
    rays_h = torch.from_numpy(scenes.incoherent_rays(n, seed=0x5EED0002).view(np.uint8).reshape(-1)).pin_memory()
    hits_h = torch.empty(n * 24, dtype=torch.uint8).pin_memory()
    rays_np, hits_np = rays_h.numpy().view(lc.Ray), hits_h.numpy().view(lc.SurfaceHit)
    rb, hb = dev.create_buffer(n, 32, 16), dev.create_buffer(n, 24, 8)
    up, run, down = dev.create_stream(), dev.create_stream(), dev.create_stream()
    ev_up, ev_run = dev.create_event(), dev.create_event()
    ext = [torch.cuda.ExternalStream(s.cuda_stream()) for s in (up, run, down)]
    serial = 0
    for rep in range(3):
        marks = []
        host = []
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        origin = torch.cuda.Event(enable_timing=True); origin.record(ext[0])
        for b0 in range(0, n, chunk):
            cnt = min(chunk, n - b0); serial += 1
            e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            h0 = time.perf_counter()
            up.submit([rb.view(b0, cnt).copy_from_async(rays_np[b0:b0 + cnt])]); e[0].record(ext[0])
            h1 = time.perf_counter()
            ev_up.signal(up, serial); ev_up.wait(run, serial)
            run.submit([shader.dispatch_async((cnt, 1, 1), rb.view(b0, cnt), hb.view(b0, cnt), accel)]); e[1].record(ext[1])
            h2 = time.perf_counter()
            ev_run.signal(run, serial); ev_run.wait(down, serial)
            down.submit([hb.view(b0, cnt).copy_to_async(hits_np[b0:b0 + cnt])]); e[2].record(ext[2])
            h3 = time.perf_counter()
            marks.append(e); host.append((h1 - h0, h2 - h1, h3 - h2))
        down.synchronize()
        total = time.perf_counter() - t0
        print(f"rep {rep}: total {total * 1e3:.2f} ms = {n / total / 1e6:.0f} Mrays/s, chunk {chunk}")
        for c, (e, h) in enumerate(zip(marks, host)):
            print(f"  chunk {c}: upload done @{origin.elapsed_time(e[0]):7.2f}  kernel done @{origin.elapsed_time(e[1]):7.2f}  download done @{origin.elapsed_time(e[2]):7.2f} ms | host in submit: up {h[0] * 1e3:.2f} run {h[1] * 1e3:.2f} down {h[2] * 1e3:.2f} ms")
    dev.close()


if __name__ == "__main__":
    main()
