#!/usr/bin/env python
"""bench.py — headline benchmark of the B200 ray-tracing device.

Metric (BASELINE.json): Mrays/s closest-hit on incoherent rays, with BVH build ms beside it, as absolute numbers and as a
fraction of the measured HBM roofline.  Workload at every N: C3 of BASELINE.json — a 1 000 000-triangle random soup and
16 777 216 incoherent rays per GPU (`configs[2]`, the configuration the metric is quoted on).  A step = one pass of the reference
call path over the ray batch: the DSL kernel `hits.write(i, accel.intersect(rays.read(i), mask))` (rtx.rs:774-795) as an
ir::KernelModule through DeviceInterface.create_shader + dispatch(ShaderDispatch) — lowered by the device to its persistent
wavefront form (csrc/ir_lower.cpp).  `value` times that with rays resident in HBM; `e2e` times the same kernel with pinned HOST
buffers, BufferUpload / ShaderDispatch / BufferDownload commands pipelined over three streams with timeline events — only
DeviceInterface calls.  The batch entry points (lc_b200_trace_closest, ..._host) are reported beside them as extra keys, and so are
config C2 (Cornell path tracer through create_shader) and config C5 (4K path tracing over the 10-instance 50 M-triangle scene,
tiles sharded over the ranks, render + NCCL framebuffer gather in one timed region).

  python bench.py --gpus N --steps K --warmup W          our arm (one rank per GPU under torchrun for N > 1)
  python bench.py --impl reference ...                   the CPU arm: the oracle port on all host cores, bounded sample

Scaling is weak: Accel replicated (every rank builds it from the same seeded data), each rank traces its own 16 Mi-ray
batch, no collective on the data path; hits stay on the GPU that traced them.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")   # before the CUDA context exists (luisa-compute-rs_b200/__init__.py says why)

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

N_TRIS = 1_000_000
N_RAYS = 1 << 24
METRIC = "closest_hit_incoherent_mrays_per_s"
UNIT = "Mrays/s"
WORKLOAD = "C3: 1M-triangle random soup (seed 0x5EED0001), 16Mi incoherent rays per GPU (seed 0x5EED0002+rank), closest hit through create_shader + ShaderDispatch"


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons (B200_PROFILING.md recipe).  Started before the warm-up (nvidia-smi takes a few
    hundred ms to deliver its first row) and polled every 20 ms; `stop(t0, t1)` summarises the rows whose arrival time
    falls inside the timed region [t0, t1] (host clock), falling back to the rows since warm-up began if the region was
    shorter than one polling interval."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, bufsize=1)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def wait_first(self, timeout=3.0):
        t_end = time.perf_counter() + timeout
        while self.proc and not self.rows and time.perf_counter() < t_end:
            time.sleep(0.01)

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        inside = [r for t, r in self.rows if t0 <= t <= t1 + 0.03]
        window = "timed region"
        if not inside:
            inside = [r for t, r in self.rows if t >= t0 - 1.0]
            window = "warm-up + timed region (timed region shorter than one polling interval)"
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in inside:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for name, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm),
                "window": window}


def cpu_leg(n_sample, threads=0, repeats=1):
    """The oracle port in its fast mode (multi-threaded SAH build, 8-wide tree with AVX2 box tests, the canonical triangle arithmetic —
    the same hits as the scalar checker modes, tests/test_oracle.py) on `threads` host threads over the first n_sample rays of rank 0's
    batch.  Returns Mrays/s."""
    import oracle_lib as ol
    import scenes
    desc = scenes.c3_soup(N_TRIS)
    t0 = time.perf_counter()
    o = ol.scene_from_desc(desc)
    build_s = time.perf_counter() - t0
    rays = scenes.incoherent_rays(n_sample, seed=0x5EED0002)
    best = float("inf")
    for _ in range(repeats):
        t0 = time.perf_counter()
        o.trace_closest(rays, 0xFF, ol.WIDE, threads)
        best = min(best, time.perf_counter() - t0)
    o.close()
    return n_sample / best / 1e6, build_s, best


def run_reference(args, rank):
    if rank != 0:
        return
    import oracle_lib as ol
    cores = ol.lib().oracle_hw_threads()
    n_sample = 1 << 22
    import scenes
    desc = scenes.c3_soup(N_TRIS)
    o = ol.scene_from_desc(desc)
    rays = scenes.incoherent_rays(n_sample, seed=0x5EED0002)
    for _ in range(args.warmup):
        o.trace_closest(rays[: n_sample // 8], 0xFF, ol.WIDE, 0)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        o.trace_closest(rays, 0xFF, ol.WIDE, 0)
    dt = (time.perf_counter() - t0) / args.steps
    v = n_sample / dt / 1e6
    sample = f"first {n_sample} rays of the C3 batch per step against the full 1M-triangle scene; oracle port, fast mode: binned-SAH BVH collapsed 8-wide, AVX2 box tests, eight rays in flight per thread, canonical fp32 triangle test, all host threads (not Embree: the reference's own arithmetic cannot be built here)"
    emit(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": sample},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


NCU_SUMMARY = "profiles/r02_c3_lc_kernel_ncu_full_summary.csv"


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of the timed kernel (the wavefront-lowered DSL kernel, `lc_kernel`) per launch on
    this workload, from the committed `ncu --set full` capture of `bench.py --profile` (bytes); None if the summary is not there."""
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    try:
        tot = 0.0
        for line in open(os.path.join(ROOT, NCU_SUMMARY)):
            f = line.strip().split(",")
            if len(f) == 3 and f[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                tot += float(f[2]) * scale[f[1]]
        return tot or None
    except OSError:
        return None


def dsl_path_tracer_leg(dev, lc, scenes, torch, _ext):
    """Cornell box 1024 x 1024, 32 spp per dispatch, depth 10 (BASELINE configs[1]) with the example's kernel built as an
    ir::KernelModule and lowered by the device; rays counted by the hand-lowered twin on the same seeds (bit-identical walks)."""
    import ctypes as C
    import luisa_compute_rs_b200.examples as ex
    from luisa_compute_rs_b200 import examples_ir
    w = h = 1024
    desc = scenes.c2_cornell()
    pt = ex.PathTracer(dev, desc.meshes, w, h)
    for _ in range(2):
        pt.dispatch(32, 10, count_rays=True)
    rays_per_dispatch = sum(pt.rays) / 2
    n = len(desc.meshes)
    vheap, iheap = dev.create_bindless_array(n), dev.create_bindless_array(n)
    for i, (vb, ib) in enumerate(zip(pt.vbuffers, pt.ibuffers)):
        vheap.emplace_buffer_async(i, vb); iheap.emplace_buffer_async(i, ib)
    s = dev.default_stream()
    s.submit([vheap.update_async(), iheap.update_async()])
    image = dev.create_tex2d("Rgba32f", w, h); seeds = dev.create_tex2d("R32Uint", w, h)
    seeds.copy_from(ex.seed_image(w, h).reshape(h, w))
    t0 = time.perf_counter()
    # as the example runs it: libdevice sin / cos and the frontend's default build options (enable_fast_math = true, runtime/kernel.rs:564)
    k = examples_ir.path_tracer_kernel(vheap.handle.id, iheap.handle.id, 32, 10, polynomial_sincos=False)
    shader = dev.create_shader(C.addressof(k.km), fast_math=True, keep=k)
    create_s = time.perf_counter() - t0
    res = np.array([w, h], np.uint32)
    shader.dispatch((w, h), image, seeds, pt.accel, res)   # warm-up; same seeds as the counted dispatches above
    sext = torch.cuda.ExternalStream(s.cuda_stream())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 4
    e0.record(sext)
    s.submit([shader.dispatch_async((w, h), image, seeds, pt.accel, res) for _ in range(reps)])
    e1.record(sext); s.synchronize()
    ms = e0.elapsed_time(e1) / reps
    out = {"workload": "C2: Cornell box 1024x1024, 32 spp per dispatch, depth 10, examples/path_tracer.rs as IR through create_shader (enable_fast_math = the frontend default)",
           "ms_per_dispatch": ms, "mrays_per_s": rays_per_dispatch / ms / 1e3, "rays_per_dispatch": rays_per_dispatch, "create_shader_s": create_s,
           "note": "ray count of the first two dispatches of the hand-lowered twin; later dispatches trace a similar number"}
    for r in (shader, image, seeds, vheap, iheap):
        r.destroy()
    pt.destroy()
    return out


def bind_to_gpu_numa_node(torch, local_rank):
    """Run this rank — and allocate its pinned staging — on the host cores next to its GPU: with 8 ranks the e2e legs move 56 B per ray
    through host memory, and a rank whose pinned buffers sit on the other socket crosses the socket interconnect twice.  No-op when the
    platform reports no NUMA node for the device (single-socket hosts, containers that hide /sys)."""
    info = {"numa_node": None, "cpus": None}
    try:
        import ctypes
        bdf = None
        buf = ctypes.create_string_buffer(32)
        rt = ctypes.CDLL("libcudart.so.12")
        if rt.cudaDeviceGetPCIBusId(buf, 32, local_rank) == 0:
            bdf = buf.value.decode()   # "0000:1b:00.0" (torch's pci_bus_id property is the bus number alone)
        if not bdf:
            return info
        path = f"/sys/bus/pci/devices/{str(bdf).lower()}/numa_node"
        node = int(open(path).read().strip())
        info["numa_node"] = node
        if node < 0:
            return info
        cpulist = open(f"/sys/devices/system/node/node{node}/cpulist").read().strip()
        cpus = set()
        for part in cpulist.split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = cpus & set(os.sched_getaffinity(0))
        if allowed:
            os.sched_setaffinity(0, allowed)
            info["cpus"] = len(allowed)
    except Exception as e:  # binding is an optimisation: never fail the bench for it
        info["error"] = str(e)[:80]
    return info


def pcie_probe(torch, dist, world, mib=512, reps=3):
    """What the host link gives THIS run: every rank copies `mib` MiB pinned -> device and device -> pinned at the same time, all ranks
    together; returns per-rank GB/s (min over ranks) and the sum — the ceiling of any host-buffer (e2e) number at this N."""
    n = mib << 20
    h_in = torch.empty(n, dtype=torch.uint8).pin_memory(); h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
    d_in = torch.empty(n, dtype=torch.uint8, device="cuda"); d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    best = float("inf")
    for _ in range(reps):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        with torch.cuda.stream(s1):
            d_in.copy_(h_in, non_blocking=True)
        with torch.cuda.stream(s2):
            h_out.copy_(d_out, non_blocking=True)
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    gbs = torch.tensor([2 * n / best / 1e9], device="cuda", dtype=torch.float64)
    lo, total = gbs.clone(), gbs.clone()
    if world > 1:
        dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(total, op=dist.ReduceOp.SUM)
    return {"h2d_plus_d2h_gbs_per_rank_min": float(lo.item()), "h2d_plus_d2h_gbs_sum": float(total.item()),
            "e2e_ceiling_mrays_per_s": float(total.item()) * 1e9 / 56 / 1e6, "note": f"{mib} MiB each way per rank, concurrently on all {world} ranks, pinned memory"}



def e2e_reference_api(dev, shader, accel, rb, hb, rays_np, hits_np, n, lanes, chunk_rays):
    """One end-to-end pass with HOST buffers through DeviceInterface calls only: per chunk a BufferUpload on the upload stream, the
    ShaderDispatch over that chunk's buffer views on the compute stream, a BufferDownload on the download stream, ordered by two
    timeline events — what a luisa-compute-rs program does with three Streams and two Events.  One host thread: a BufferUpload only
    returns once its source has been read (the snapshot contract, cpu/stream.rs:33-64), so the link idles for the ~0.25 ms the other
    five calls of a chunk take.  (A second Python thread for those calls was measured and is slower — 543 Mrays/s against 750-910: the
    GIL changes hands at millisecond granularity; profiles/r02s_bench.json.)"""
    up, run, down, ev_up, ev_run = lanes
    base = e2e_reference_api.serial
    c = 0
    for b0 in range(0, n, chunk_rays):
        cnt = min(chunk_rays, n - b0)
        c += 1
        up.submit([rb.view(b0, cnt).copy_from_async(rays_np[b0:b0 + cnt])])
        ev_up.signal(up, base + c); ev_up.wait(run, base + c)
        run.submit([shader.dispatch_async((cnt, 1, 1), rb.view(b0, cnt), hb.view(b0, cnt), accel)])
        ev_run.signal(run, base + c); ev_run.wait(down, base + c)
        down.submit([hb.view(b0, cnt).copy_to_async(hits_np[b0:b0 + cnt])])
    e2e_reference_api.serial = base + c
    down.synchronize()


e2e_reference_api.serial = 0


def parity_sample_leg(dev, lc, shader_factory, n_sample=1 << 20):
    """north_star's tolerance statement made driver-visible: on the first 1 Mi rays of the batch against the full 1 M-triangle scene,
    hits of the DSL call path vs the oracle's float64 ground truth — inst / prim must agree wherever the f64 answer is unambiguous,
    t and barycentrics within 1e-5; the rate of ambiguous rays (a second candidate within 1e-5 relative, or an edge within rounding)
    is the tie rate."""
    import oracle_lib as ol
    import scenes
    from harness import DeviceScene, compare_with_truth
    desc = scenes.c3_soup(N_TRIS)
    rays = scenes.incoherent_rays(n_sample, seed=0x5EED0002)
    d = DeviceScene(dev, desc)
    got = d.trace_dsl(rays)
    o = ol.scene_from_desc(desc)
    t0 = time.perf_counter()
    truth, amb = o.truth(rays)
    truth_s = time.perf_counter() - t0
    r = compare_with_truth(got, truth, amb)
    canon = o.trace_closest(rays[: 1 << 16])
    bit_equal = bool(got[: 1 << 16].tobytes() == canon.tobytes())
    d.destroy(); o.close()
    return {"rays": n_sample, "tie_rate": r["tie_rate"], "f64_mismatches": r["mismatches"], "disagree_inside_ties": r["disagree_in_ties"], "t_rel_err_max": r["t_rel_err"],
            "bary_abs_err_max": r["bary_abs_err"], "bit_identical_to_oracle_on_64k": bit_equal, "truth_seconds": round(truth_s, 2),
            "note": "DSL call path (create_shader + ShaderDispatch) vs oracle float64 ground truth; mismatches counted outside flagged ties"}


def c5_leg(dev, lc, scenes, torch, dist, rank, world, spp, spp_per_dispatch, balance_passes):
    """BASELINE config C5 (4K, depth 5, `spp` samples per pixel, 10 x 5 M-triangle instances + light): render AND the NCCL framebuffer
    gather inside one timed region on the device, strong scaling (the frame is fixed, ranks split it)."""
    from luisa_compute_rs_b200.tiled_render import TiledPathTracer
    t0 = time.perf_counter()
    pt = TiledPathTracer(dev, lc, scenes, spp_per_dispatch=spp_per_dispatch, rank=rank, world=world, dist=dist)
    setup_s = time.perf_counter() - t0
    pt.frame(spp_per_dispatch, first_frame=50000)                  # warm-up: kernels, NCCL communicator, allocator
    if world > 1:
        pt.cost_from_ray_counts()   # per-tile ray counts of one short pass: the cost map at tile resolution
    history = pt.balance(balance_passes) if world > 1 else []
    recuts = 0   # in-frame re-cuts (tiled_render.frame) are implemented and image-exact, but measured neutral at N = 2 / 4 and 2 % slower at N = 8 (profiles/r02u_c5_n*.jsonl)
    ms, gathered, n_dispatch = pt.frame(spp, first_frame=0, recuts=recuts)
    times = pt.all_times(ms)
    rays = pt.counters_t[:2].clone()
    if world > 1:
        dist.all_reduce(rays, op=dist.ReduceOp.SUM)
    img = pt.image(gathered) if rank == 0 else None
    render_only = pt.all_times(pt.timed(lambda: pt.render(1, 70000)))   # one more pass without the gather, for the per-rank picture
    out = None
    if rank == 0:
        total = max(times)
        imbalance = max(render_only) / (sum(render_only) / len(render_only))
        out = {"workload": "C5: 10 instances of a 5.0M-triangle terrain + emissive quad (50.0M triangles), 3840x2160, depth 5, path tracer as IR through create_shader; "
                           "tiles sharded over the ranks, Accel replicated, ONE ncclAllGather of the framebuffer inside the timed region",
               "n_gpus": world, "spp": n_dispatch * spp_per_dispatch, "spp_per_dispatch": spp_per_dispatch, "triangles": pt.triangles,
               "frame_ms": total, "frame_ms_per_rank": [round(t, 2) for t in times], "scaling": "strong",
               "mrays_per_s": float(rays.sum().item()) / total / 1e3,
               "msamples_per_s": pt.width * pt.height * n_dispatch * spp_per_dispatch / total / 1e3,
               "partition": "contiguous ranges of the Morton-ordered 64x64 tiles, cut to equal measured cost and re-cut inside the frame (tiles that change owner take their accumulators along)" if world > 1 else "all tiles on one GPU",
               "recuts_in_frame": [{"imbalance_before": r[0], "tiles_moved_by_rank0": r[1]} for r in getattr(pt, "recut_log", [])],
               "tiles_per_rank": [int(x) for x in np.diff(pt.bounds)], "balance_passes": [{"imbalance": round(h[0], 3), "tiles_per_rank": h[1], "ms_per_rank": h[2]} for h in history],
               "one_pass_ms_per_rank": [round(t, 3) for t in render_only], "imbalance_max_over_mean": round(imbalance, 4),
               "limiter": ("load imbalance between ranks" if imbalance > 1.08 else "per-rank efficiency: the ranks that own the expensive tiles own few of them (70-160 tiles = 3-6 waves of resident threads per dispatch), so dispatch tails weigh more than on one GPU"),
               "gather_bytes": int(gathered.numel() * 4) if world > 1 else 0, "blas_build_ms": round(pt.blas_ms, 3), "tlas_build_ms": round(pt.tlas_ms, 3),
               "spp_per_pixel_ok": bool(np.all(img[..., 3] == n_dispatch)), "image_sha256": pt.sha(img), "setup_s": round(setup_s, 1)}
    pt.destroy()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--tris", type=int, default=N_TRIS)
    ap.add_argument("--rays", type=int, default=N_RAYS)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--profile", action="store_true", help="short run for ncu: only the headline steps")
    ap.add_argument("--c5-spp", type=int, default=1024, help="samples per pixel of the C5 leg (BASELINE: 1024); 0 skips the leg")
    ap.add_argument("--c5-spp-per-dispatch", type=int, default=32, help="the example's own SPP_PER_DISPATCH (path_tracer.rs:179)")
    ap.add_argument("--e2e-chunk", type=int, default=1 << 21, help="rays per chunk of the host-buffer pipeline")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    warmup = max(args.warmup, 3) if not args.profile else args.warmup

    import ctypes as C
    import torch
    import torch.distributed as dist
    import luisa_compute_rs_b200 as lc
    import scenes
    from luisa_compute_rs_b200 import examples_ir

    torch.cuda.set_device(local_rank)
    os.environ["LC_B200_DEVICE"] = str(local_rank)
    # before any pinned allocation; only when several ranks share the host (at N = 1 the CPU legs keep every core the box gives)
    host_binding = bind_to_gpu_numa_node(torch, local_rank) if world > 1 else {"numa_node": None, "cpus": None, "note": "not bound at N = 1"}
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    ctx = lc.Context()
    dev = ctx.create_device("b200")
    lib = lc._abi.load_library()

    # ---- scene: replicated on every rank, built by the device --------------------------------------------------
    verts, tris = scenes.random_soup(args.tris, 0x5EED0001)
    vb = dev.create_buffer_from_array(verts)
    ib = dev.create_buffer_from_array(tris)
    mesh = dev.create_mesh(vb.view(), ib.view(), lc.AccelOption())
    build_ms = []
    for _ in range(4):
        mesh.build(lc.AccelBuildRequest.FORCE_BUILD)
        build_ms.append(mesh.stats()["build_ms"])
    mstats = mesh.stats()
    accel = dev.create_accel()
    accel.push_mesh(mesh)
    tlas_all = []
    for _ in range(3):
        accel.build()
        tlas_all.append(accel.stats()["build_ms"])
    tlas_ms = min(tlas_all)

    # ---- the DSL kernel of the reference call path, lowered + compiled by the device ------------------------------------
    t0 = time.perf_counter()
    kernel = examples_ir.trace_buffer_kernel()
    shader = dev.create_shader(C.addressof(kernel.km), keep=kernel)
    create_shader_s = time.perf_counter() - t0

    # ---- rays: resident in HBM before the timed region -------------------------------------------------------------
    n = args.rays
    rays_h = torch.from_numpy(scenes.incoherent_rays(n, seed=0x5EED0002 + rank).view(np.uint8).reshape(-1)).pin_memory()
    hits_h = torch.empty(n * 24, dtype=torch.uint8).pin_memory()
    rays_np, hits_np = rays_h.numpy().view(lc.Ray), hits_h.numpy().view(lc.SurfaceHit)
    rb = dev.create_buffer(n, 32, 16)
    hb = dev.create_buffer(n, 24, 8)
    ob = dev.create_buffer(n, 4, 4)
    rb.view().copy_from(rays_np)
    stream = dev.create_stream()
    ext = torch.cuda.ExternalStream(stream.cuda_stream(), device=torch.device("cuda", local_rank))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        stream.submit([shader.dispatch_async((n, 1, 1), rb, hb, accel)])

    sampler = ClockSampler(local_rank)
    sampler.start()
    sampler.wait_first()
    for _ in range(warmup):
        step()
    stream.synchronize()

    barrier()
    t_region0 = time.perf_counter()
    launches0 = lib.lc_b200_kernel_launch_count()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    evs[0].record(ext)
    for k in range(args.steps):
        step()
        evs[k + 1].record(ext)
    stream.synchronize()
    barrier()
    t_region1 = time.perf_counter()
    launches = lib.lc_b200_kernel_launch_count() - launches0
    clocks = sampler.stop(t_region0, t_region1)
    step_ms = [evs[k].elapsed_time(evs[k + 1]) for k in range(args.steps)]
    total_ms = float(sum(step_ms))
    if world > 1:
        t = torch.tensor([total_ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    value = world * n * args.steps / (total_ms * 1e-3) / 1e6
    kernel_ms = float(np.mean(step_ms))

    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warmup,
        "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": WORKLOAD if (args.tris, args.rays) == (N_TRIS, N_RAYS) else f"C3-shaped: {args.tris} triangles, {args.rays} rays per GPU",
                   "triangles": args.tris, "rays_per_gpu": n, "parallelism": f"rays sharded x{world}, accel replicated",
                   "api": "DeviceInterface.create_shader(ir::KernelModule of `hits.write(i, accel.intersect(rays.read(i), 0xff))`) + dispatch(ShaderDispatch)",
                   "lowering": "wavefront (persistent threads, trace calls are suspension points of the warp-synchronous traversal loop)",
                   "l2": "ray + hit buffers (896 MiB) exceed the 126 MB L2; the BVH (~80 MB) is L2-resident by nature of the workload"},
        "gpu_launches": int(launches), "clocks": clocks, "create_shader_s": create_shader_s,
        "build": {"blas_ms": float(min(build_ms)), "blas_ms_all": [float(x) for x in build_ms], "tlas_ms": float(tlas_ms),
                  "wide_nodes": int(mstats["wide_node_count"]), "bvh_bytes": int(mstats["bvh_bytes"]), "max_depth": int(mstats["max_depth"]),
                  "builder": ["lbvh", "ploc"][int(mstats["builder"])] + " (chosen per mesh under AccelUsageHint::FastTrace)"},
    }

    # the batch entry point on the same buffers (the native extension symbol, not the reference call path): every rank, device-timed
    for _ in range(2):
        accel.intersect(rb, hb, n, 0xFF, stream)
    stream.synchronize()
    dsl_hits = None
    if rank == 0 and not args.profile:
        step(); stream.synchronize()
        dsl_hits = np.empty(n, dtype=lc.SurfaceHit); hb.view().copy_to(dsl_hits)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(ext)
    for _ in range(5):
        accel.intersect(rb, hb, n, 0xFF, stream)
    e1.record(ext); stream.synchronize()
    batch_ms = e0.elapsed_time(e1) / 5
    out["batch_entry"] = {"value": n / batch_ms / 1e3, "unit": UNIT, "ms": batch_ms, "api": "lc_b200_trace_closest (k_trace + k_refine); per GPU", "dsl_over_batch": batch_ms / kernel_ms}

    # ---- config C5 on all ranks, before the legs only rank 0 runs (they would leave its GPU warmer than the others) ----
    if not args.profile and args.c5_spp > 0:
        c5 = c5_leg(dev, lc, scenes, torch, dist if world > 1 else None, rank, world, args.c5_spp, args.c5_spp_per_dispatch, 6)
        if rank == 0:
            out["c5_path_trace"] = c5

    if rank == 0 and not args.profile:
        batch_hits = np.empty(n, dtype=lc.SurfaceHit); hb.view().copy_to(batch_hits)
        assert dsl_hits.tobytes() == batch_hits.tobytes(), "the DSL call path and the batch entry point disagree"
        # ---- roofline of the timed kernel: algorithmic bytes from the instrumented traversal on the same inputs (the per-ray walk is
        #      the same in the batch kernel and in the lowered kernel: it depends on the ray and the tree only) ----
        ctr = accel.intersect_counted(rb, hb, n, 0xFF, stream)
        algo_bytes = n * (32 + 24) + 128 * ctr["nodes_visited"] + 48 * ctr["tris_tested"]
        peak, how = measured_peaks()
        achieved = algo_bytes / (kernel_ms * 1e-3) / 1e9
        build_bytes = 12 * args.tris + 12 * verts.shape[0] + mstats["bvh_bytes"]
        traffic = ncu_traffic()
        out["roofline"] = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                           "traffic_source": NCU_SUMMARY, "peak_source": how, "kernel": "lc_kernel (wavefront-lowered DSL kernel: traversal + barycentrics in one launch)",
                           "algorithmic_bytes_per_launch": int(algo_bytes), "nodes_per_ray": ctr["nodes_visited"] / n, "tris_per_ray": ctr["tris_tested"] / n,
                           "dram_frac": (traffic / (kernel_ms * 1e-3) / 1e9 / peak) if traffic else None,
                           "compulsory_dram_bytes": int(n * 56 + mstats["bvh_bytes"]),
                           "note": "frac is the LOGICAL (L1/L2-inclusive) fraction of SURVEY.md §8(d): 56 B/ray I/O + 128 B per node visit + 48 B per triangle test; it grows with nodes_per_ray, "
                                   "so Mrays/s and nodes_per_ray are the co-metrics.  dram_frac is measured DRAM traffic (ncu) over the same time: the kernel is bound by L1TEX wavefronts of divergent "
                                   "32-byte gathers + issue slots, not by HBM",
                           "build_achieved_gbs": build_bytes / (min(build_ms) * 1e-3) / 1e9, "build_frac": build_bytes / (min(build_ms) * 1e-3) / 1e9 / peak}
        # ---- any-hit on the same batch (C3's shadow set), both call paths ----
        shadow = scenes.shadow_rays_from_hits(rays_np, dsl_hits)
        rb2 = dev.create_buffer_from_array(shadow)
        ka = examples_ir.trace_buffer_kernel(any_hit=True)
        sha = dev.create_shader(C.addressof(ka.km), keep=ka)
        res = {}
        for name, fn in (("dsl", lambda: stream.submit([sha.dispatch_async((n, 1, 1), rb2, ob, accel)])), ("batch", lambda: accel.intersect_any(rb2, ob, n, 0xFF, stream))):
            for _ in range(2):
                fn()
            stream.synchronize()
            e0.record(ext)
            for _ in range(5):
                fn()
            e1.record(ext); stream.synchronize()
            res[name] = n * 5 / (e0.elapsed_time(e1) * 1e-3) / 1e6
        out["any_hit"] = {"value": res["dsl"], "unit": UNIT, "batch_entry": res["batch"], "hit_rate": float((dsl_hits["inst"] != lc.INVALID).mean())}
        sha.destroy(); rb2.destroy()

    if not args.profile:
        # ---- end to end with HOST buffers through DeviceInterface calls only (docstring) ----
        lanes = (dev.create_stream(), dev.create_stream(), dev.create_stream(), dev.create_event(), dev.create_event())
        e2e_reference_api(dev, shader, accel, rb, hb, rays_np, hits_np, n, lanes, args.e2e_chunk)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_reference_api(dev, shader, accel, rb, hb, rays_np, hits_np, n, lanes, args.e2e_chunk)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        out["e2e"] = {"value": world * n * args.steps / dt / 1e6, "unit": UNIT, "h2d_bytes_per_step": world * n * 32, "d2h_bytes_per_step": world * n * 24,
                      "api": f"DeviceInterface only: per {args.e2e_chunk}-ray chunk BufferUpload | ShaderDispatch | BufferDownload on three streams ordered by timeline events; pinned host buffers"}
        if dsl_hits is not None:
            assert hits_np.tobytes() == dsl_hits.tobytes(), "host-buffer pipeline and device-resident dispatch disagree"
        # the batch-form host entry point beside it
        accel.intersect_host_ptr(rays_h.data_ptr(), hits_h.data_ptr(), n)
        barrier()
        t0 = time.perf_counter()
        for _ in range(max(args.steps // 2, 1)):
            accel.intersect_host_ptr(rays_h.data_ptr(), hits_h.data_ptr(), n)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        out["e2e_batch_entry"] = {"value": world * n * max(args.steps // 2, 1) / dt / 1e6, "unit": UNIT, "api": "lc_b200_trace_closest_host (chunked H2D / k_trace / D2H pipeline inside the library)"}
        for r in lanes:
            r.destroy()
        out["host_link"] = pcie_probe(torch, dist if world > 1 else None, world)
        out["host_link"]["rank0_binding"] = host_binding

    if rank == 0 and not args.profile:
        out["parity_sample"] = parity_sample_leg(dev, lc, None)
        # ---- config C2 beside the headline: examples/path_tracer.rs as an IR kernel through create_shader ----
        out["dsl_path_tracer"] = dsl_path_tracer_leg(dev, lc, scenes, torch, ext)

    if rank == 0 and world == 1 and not args.no_cpu and not args.profile:
        import oracle_lib as ol
        cores = ol.lib().oracle_hw_threads()
        n_sample = min(n, 1 << 24)   # ~7 s on 16 cores: the whole batch
        v, cpu_build_s, cpu_s = cpu_leg(n_sample)
        out["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                               "sample": f"first {n_sample} rays of the batch ({cpu_s:.1f} s) on the full 1M-triangle scene; oracle port, fast mode (SAH BVH built in {cpu_build_s:.1f} s on all threads, collapsed 8-wide, AVX2 box tests, eight rays in flight per thread, canonical triangle test), not Embree"}

    if rank == 0:
        emit(json.dumps(out))
    for b in (shader, rb, hb, ob, vb, ib):
        b.destroy()
    accel.destroy(); mesh.destroy(); stream.destroy(); dev.close()
    if world > 1:
        dist.destroy_process_group()


def _claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries below us write there too (NCCL prints its version banner on stdout at
    any NCCL_DEBUG level >= VERSION): keep the real stdout aside for the final line and point fd 1 at stderr for everything else."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def emit(line):
    os.write(_REAL_STDOUT, (line + "\n").encode())


_REAL_STDOUT = 1

if __name__ == "__main__":
    _claim_stdout()
    main()
