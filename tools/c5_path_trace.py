#!/usr/bin/env python
"""BASELINE config C5 stand-alone: path tracing over the 10-instance 50M-triangle scene with image tiles sharded across GPUs
(`python -m torch.distributed.run --nproc-per-node N tools/c5_path_trace.py ...`, or plain python for N = 1).  The host program is
luisa-compute-rs_b200/tiled_render.py (the same one `bench.py`'s c5_path_trace leg runs); this tool exposes its knobs for sweeps.
Prints one JSON line from rank 0."""
import argparse
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
    ap.add_argument("--spp-per-dispatch", type=int, default=32)
    ap.add_argument("--depth", type=int, default=5)
    ap.add_argument("--nx", type=int, default=1582, help="terrain vertices per side (1582 -> 5.0M triangles per mesh, 50M instanced)")
    ap.add_argument("--block", type=int, default=8, help="edge of the kernel's square thread block")
    ap.add_argument("--streams", type=int, default=2)
    ap.add_argument("--balance-passes", type=int, default=6, help="0: equal tile counts per rank")
    ap.add_argument("--recuts", type=int, default=0, help="re-cuts of the ranges inside the frame (world > 1)")
    ap.add_argument("--also-recuts", type=int, default=-1, help="render the frame a second time with this many re-cuts (same process: A/B without a second set-up)")
    ap.add_argument("--emulate", default="", help="RANK/WORLD: render only that rank's equal-count range in this single process (tuning aid)")
    ap.add_argument("--precise", action="store_true", help="compile the kernel without enable_fast_math (the frontend default is on)")
    ap.add_argument("--lowering", default="auto", choices=["auto", "direct", "wavefront"])
    ap.add_argument("--save", default="")
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local_rank)
    os.environ["LC_B200_DEVICE"] = str(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    import luisa_compute_rs_b200 as lc
    import scenes
    from luisa_compute_rs_b200 import sharding
    from luisa_compute_rs_b200.tiled_render import TiledPathTracer
    dev = lc.Context().create_device("b200")
    lc._abi.load_library().lc_b200_set_lowering({"auto": 0, "direct": 1, "wavefront": 2}[a.lowering])
    erank, eworld = (int(v) for v in a.emulate.split("/")) if a.emulate else (rank, world)
    pt = TiledPathTracer(dev, lc, scenes, a.width, a.height, a.nx, a.spp_per_dispatch, a.depth, a.block, a.streams, erank, eworld, dist if world > 1 else None, fast_math=not a.precise)
    if a.emulate:
        pt.world = 1   # no collective: time this rank's share alone
        pt.frame(a.spp_per_dispatch, 50000)
        ms = pt.timed(lambda: pt.render(max(1, a.spp // a.spp_per_dispatch), 0))
        os.write(real_stdout, (json.dumps({"emulate": a.emulate, "tiles": pt.b1 - pt.b0, "render_ms": round(ms, 3), "spp": a.spp}) + "\n").encode())
        dev.close()
        return
    pt.frame(a.spp_per_dispatch, 50000)
    if world > 1 and a.balance_passes:
        pt.cost_from_ray_counts()   # per-tile ray counts of one short pass: the cost map at tile resolution
    history = pt.balance(a.balance_passes) if world > 1 and a.balance_passes else []
    ms, gathered, n_dispatch = pt.frame(a.spp, 0, recuts=a.recuts if world > 1 else 0)
    times = pt.all_times(ms)
    rays = pt.counters_t[:2].clone()
    if world > 1:
        dist.all_reduce(rays, op=dist.ReduceOp.SUM)
    if rank == 0:
        img = pt.image(gathered)
        total_rays = int(rays.sum().item())
        rgb = img[..., :3] / np.maximum(img[..., 3:4], 1)
        res = {"config": "c5_path_trace", "n_gpus": world, "width": a.width, "height": a.height, "spp": n_dispatch * a.spp_per_dispatch, "depth": a.depth, "triangles": pt.triangles,
               "frame_ms": round(max(times), 3), "frame_ms_per_rank": [round(t, 2) for t in times], "streams": len(pt.lanes), "block": a.block, "fast_math": not a.precise, "lowering": a.lowering,
               "tiles_per_rank": [int(x) for x in np.diff(pt.bounds)], "balance_passes": [{"imbalance": round(h[0], 3), "tiles_per_rank": h[1], "ms_per_rank": h[2]} for h in history], "rays": total_rays,
               "mrays_per_s": round(total_rays / max(times) / 1e3, 1), "mean_radiance": round(float(rgb.mean()), 5),
               "recuts_in_frame": [{"imbalance_before": r[0], "tiles_moved_by_rank0": r[1]} for r in getattr(pt, "recut_log", [])],
               "spp_per_pixel_ok": bool(np.all(img[..., 3] == n_dispatch)), "image_sha256": pt.sha(img)}
        os.write(real_stdout, (json.dumps(res) + "\n").encode())
    if a.also_recuts >= 0 and world > 1:
        ms2, g2, _ = pt.frame(a.spp, 0, recuts=a.also_recuts)
        t2 = pt.all_times(ms2)
        if rank == 0:
            img2 = pt.image(g2)
            os.write(real_stdout, (json.dumps({"config": "c5_path_trace second frame", "recuts": a.also_recuts, "frame_ms": round(max(t2), 3), "tiles_per_rank": [int(x) for x in np.diff(pt.bounds)],
                                               "recuts_in_frame": [{"imbalance_before": r[0], "tiles_moved_by_rank0": r[1]} for r in pt.recut_log], "image_sha256": pt.sha(img2)}) + "\n").encode())
    if rank == 0:
        if a.save:
            np.save(a.save, rgb.astype(np.float32))
    if world > 1:
        dist.barrier(); dist.destroy_process_group()
    dev.close()


if __name__ == "__main__":
    main()
