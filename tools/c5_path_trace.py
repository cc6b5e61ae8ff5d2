#!/usr/bin/env python
"""BASELINE config C5: path tracing over a 10-instance 50M-triangle scene with image tiles sharded across GPUs.

One process per GPU (`python -m torch.distributed.run --nproc-per-node N tools/c5_path_trace.py ...`, or plain python for N = 1).
Every rank builds the same meshes + Accel itself (replicated, no broadcast), renders the 64x64 tiles the Morton round-robin
assigns to it with the IR-lowered path tracer (examples_ir.tiled_path_tracer_kernel: create_shader + ShaderDispatch), and the
packed tile buffers are assembled by ONE NCCL all-gather — the path's only collective.  The random streams are
keyed by global pixel index, so the gathered image is bit-identical for every N; rank 0 prints its SHA-256 so runs can be compared.
Timing: CUDA events on the device's stream around the dispatches, max over ranks; the gather is timed separately."""
import argparse
import ctypes as C
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402


def main():
    real_stdout = os.dup(1); os.dup2(2, 1)   # NCCL prints its banner on stdout; keep it for the one JSON line
    ap = argparse.ArgumentParser()
    ap.add_argument("--width", type=int, default=3840)
    ap.add_argument("--height", type=int, default=2160)
    ap.add_argument("--spp", type=int, default=64, help="total samples per pixel (BASELINE: 1024)")
    ap.add_argument("--spp-per-dispatch", type=int, default=16, help="16 x 4 streams: 7.16x at N = 8; 32 x 2 streams is 2 %% faster on one GPU (profiles/r01z_c5_path_trace_16x4.jsonl)")
    ap.add_argument("--depth", type=int, default=5)
    ap.add_argument("--nx", type=int, default=1582, help="terrain vertices per side (1582 -> 5.0M triangles per mesh, 50M instanced)")
    ap.add_argument("--block", type=int, default=8, help="edge of the kernel's square thread block")
    ap.add_argument("--streams", type=int, default=4, help="the rank's tiles are split over this many streams so that the tail of one dispatch overlaps the next")
    ap.add_argument("--chunk", type=int, default=1, help="tiles per round-robin run along the Morton curve (sharding.tiles_of_rank)")
    ap.add_argument("--emulate", default="", help="RANK/WORLD: render only that rank's tiles in this single process (tuning aid; no gather)")
    ap.add_argument("--regenerate", action="store_true", help="path regeneration instead of the example's nested sample / bounce loops (same image; measured slower, profiles/r01v)")
    ap.add_argument("--save", default="")
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    import luisa_compute_rs_b200 as lc
    from luisa_compute_rs_b200 import examples_ir, sharding
    import scenes
    dev = lc.Context().create_device("b200")
    s = dev.default_stream()
    ext = torch.cuda.ExternalStream(s.cuda_stream())

    # ---- scene: 10 terrain instances on a 5 x 2 grid (yaw 36 deg * k) + one emissive quad above -------------------------
    verts, tris = scenes.terrain(a.nx)
    quad_v = np.array([[0, 0, 0], [1, 0, 0], [1, 0, 1], [0, 0, 1]], np.float32)
    quad_t = np.array([[0, 1, 2], [0, 2, 3]], np.uint32)
    vb, ib = dev.create_buffer_from_array(verts), dev.create_buffer_from_array(tris)
    qvb, qib = dev.create_buffer_from_array(quad_v), dev.create_buffer_from_array(quad_t)
    mesh = dev.create_mesh(vb.view(), ib.view(), lc.AccelOption())
    quad = dev.create_mesh(qvb.view(), qib.view(), lc.AccelOption())
    mesh.build(lc.AccelBuildRequest.FORCE_BUILD); quad.build(lc.AccelBuildRequest.FORCE_BUILD)
    mesh.build(lc.AccelBuildRequest.FORCE_BUILD)
    blas_ms = mesh.stats()["build_ms"]
    accel = dev.create_accel(lc.AccelOption())
    for k in range(10):
        t = np.eye(4, dtype=np.float32); t[:3, :] = scenes.rotation_y(36.0 * k)
        t[:3, 3] = [1.2 * (k % 5), 0.0, 1.2 * (k // 5)]
        accel.push_mesh(mesh, t)
    l_pos, l_u, l_v = (2.0, 1.6, 0.2), (2.0, 0.0, 0.0), (0.0, 0.0, 1.6)
    t = np.eye(4, dtype=np.float32); t[0, 0], t[2, 2] = l_u[0], l_v[2]; t[:3, 3] = l_pos
    accel.push_mesh(quad, t)
    accel.build(lc.AccelBuildRequest.FORCE_BUILD)
    n_inst = 11
    vheap, iheap = dev.create_bindless_array(n_inst), dev.create_bindless_array(n_inst)
    for i in range(n_inst):
        vheap.emplace_buffer_async(i, vb if i < 10 else qvb); iheap.emplace_buffer_async(i, ib if i < 10 else qib)
    s.submit([vheap.update_async(), iheap.update_async()])

    cam_o, cam_at = np.float32([3.0, 2.5, -3.0]), np.float32([3.0, 0.0, 1.0])
    f = cam_at - cam_o; f /= np.linalg.norm(f)
    r = np.cross(f, np.float32([0, 1, 0])); r /= np.linalg.norm(r)
    u = np.cross(r, f)
    camera = (tuple(map(float, cam_o)), tuple(map(float, f)), tuple(map(float, r)), tuple(map(float, u)), float(np.tan(np.radians(45.0) / 2)))
    light = (l_pos, l_u, l_v, (60.0, 54.0, 45.0), 10)
    kb = examples_ir.tiled_path_tracer_kernel(vheap.handle.id, iheap.handle.id, camera, light, n_inst, a.spp_per_dispatch, a.depth, block=a.block, regenerate=a.regenerate)
    shader = dev.create_shader(C.addressof(kb.km), keep=kb)

    # ---- this rank's tiles ----------------------------------------------------------------------------------------------
    tile = sharding.TILE
    if a.emulate:
        erank, eworld = (int(v) for v in a.emulate.split("/"))
        tx, ty = sharding.tiles_of_rank(a.width, a.height, erank, eworld, chunk=a.chunk)
    else:
        tx, ty = sharding.tiles_of_rank(a.width, a.height, rank, world, chunk=a.chunk)
    tiles_x = (a.width + tile - 1) // tile
    per_rank = max(sharding.padded_tile_count(a.width, a.height, world, chunk=a.chunk), tx.shape[0])
    tile_ids = dev.create_buffer_from_array((ty * tiles_x + tx).astype(np.uint32))
    out_t = torch.zeros((per_rank * tile * tile, 4), dtype=torch.float32, device="cuda")
    out = dev.wrap_device_memory(out_t.data_ptr(), per_rank * tile * tile, 16, 16)
    counters_t = torch.zeros(2, dtype=torch.int64, device="cuda")
    counters = dev.wrap_device_memory(counters_t.data_ptr(), 2, 8, 8)
    n_dispatch = max(1, a.spp // a.spp_per_dispatch)

    def params(frame):
        return np.array([a.width, a.height, frame, tx.shape[0]], np.uint32)

    # the rank's tiles in `--streams` contiguous groups, one stream each: a dispatch's tail (a few long paths) overlaps the other
    # group's work instead of idling the GPU; the groups touch disjoint slices of the tile buffer
    lanes = [s] + [dev.create_stream() for _ in range(max(1, a.streams) - 1)]
    bounds = np.linspace(0, tx.shape[0], len(lanes) + 1).astype(int)
    done = dev.create_event()

    def render(first_frame):
        for li, lane in enumerate(lanes):
            b0, b1 = int(bounds[li]), int(bounds[li + 1])
            if b1 == b0:
                continue
            lane.submit([shader.dispatch_async((tile, tile * (b1 - b0)), tile_ids.view(b0, b1 - b0), out.view(b0 * tile * tile, (b1 - b0) * tile * tile), accel,
                                               np.array([a.width, a.height, first_frame + i, b1 - b0], np.uint32), counters) for i in range(n_dispatch)])
        render.serial += 1
        for lane in lanes[1:]:   # join the side lanes into the timed stream
            done.signal(lane, render.serial * 16 + lanes.index(lane)); done.wait(s, render.serial * 16 + lanes.index(lane))
    render.serial = 0

    render(1000); s.synchronize()   # warm-up (different frames), then reset
    for lane in lanes:
        lane.synchronize()
    out_t.zero_(); counters_t.zero_(); torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(ext); render(0); e1.record(ext); s.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    rays = counters_t.clone()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1:
        gathered = torch.empty((world,) + tuple(out_t.shape), dtype=out_t.dtype, device="cuda")
        dist.all_gather_into_tensor(gathered.view(-1), out_t.view(-1))   # NCCL warm-up (communicator setup)
        torch.cuda.synchronize(); dist.barrier()
        g0.record(); dist.all_gather_into_tensor(gathered.view(-1), out_t.view(-1)); g1.record(); torch.cuda.synchronize()
        gms = torch.tensor([g0.elapsed_time(g1)], device="cuda")
        per_rank_ms = torch.empty(world, device="cuda"); dist.all_gather_into_tensor(per_rank_ms, ms)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX); dist.all_reduce(gms, op=dist.ReduceOp.MAX); dist.all_reduce(rays, op=dist.ReduceOp.SUM)
    else:
        gathered = out_t[None]; gms = torch.zeros(1); per_rank_ms = ms.clone()
    if rank == 0 and a.emulate:
        os.write(real_stdout, (json.dumps({"emulate": a.emulate, "tiles": int(tx.shape[0]), "render_ms": round(float(ms.item()), 3), "spp_per_dispatch": a.spp_per_dispatch,
                                           "block": a.block, "streams": len(lanes), "rays": int(rays.sum().item())}) + "\n").encode())
    elif rank == 0:
        img = sharding.untile(gathered.cpu().numpy(), a.width, a.height, world, chunk=a.chunk)
        rgb = img[..., :3] / np.maximum(img[..., 3:4], 1)
        total_rays = int(rays.sum().item())
        res = {"config": "c5_path_trace", "n_gpus": world, "width": a.width, "height": a.height, "spp": n_dispatch * a.spp_per_dispatch, "depth": a.depth,
               "triangles": int(tris.shape[0]) * 10 + 2, "instances": n_inst, "blas_build_ms": round(blas_ms, 3), "tlas_build_ms": round(accel.stats()["build_ms"], 3),
               "render_ms": round(float(ms.item()), 3), "render_ms_per_rank": [round(float(x), 2) for x in per_rank_ms.tolist()], "block": a.block, "streams": len(lanes), "chunk": a.chunk, "rays": total_rays, "closest_rays": int(rays[0].item()), "any_rays": int(rays[1].item()),
               "mrays_per_s": round(total_rays / float(ms.item()) / 1e3, 1), "msamples_per_s": round(a.width * a.height * n_dispatch * a.spp_per_dispatch / float(ms.item()) / 1e3, 1),
               "gather_ms": round(float(gms.item()), 3), "gather_bytes": int(gathered.numel() * 4), "mean_radiance": round(float(rgb.mean()), 5),
               "spp_per_pixel_ok": bool(np.all(img[..., 3] == n_dispatch)), "image_sha256": hashlib.sha256(img.tobytes()).hexdigest()}
        os.write(real_stdout, (json.dumps(res) + "\n").encode())
        if a.save:
            np.save(a.save, rgb.astype(np.float32))
    if world > 1:
        dist.barrier(); dist.destroy_process_group()
    dev.close()


if __name__ == "__main__":
    main()
