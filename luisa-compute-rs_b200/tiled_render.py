"""BASELINE config C5 as a host program: tile-sharded path tracing over a replicated Accel, one process per GPU.

What a luisa-compute-rs user would write on top of the device for SURVEY.md §8e: every rank builds the same meshes + Accel itself
(the build is deterministic, so nothing is broadcast), renders its share of the 64 x 64 tiles with the path tracer as an
ir::KernelModule through create_shader + ShaderDispatch (examples_ir.tiled_path_tracer_kernel — the wavefront lowering of
csrc/ir_lower.cpp), and ONE NCCL all-gather assembles the framebuffer: the only collective of the path.

Partition (sharding.py): rank r owns a contiguous range of the Morton-ordered tiles, cut so that every range costs the same.  The cost
map is what a progressive renderer gets for free — each pass's per-rank time, exchanged with one tiny all-gather — and is refined over
a few short balancing passes before the frame starts (their samples are discarded; they are part of the warm-up, not of the frame) and
keeps being refined WHILE the frame accumulates: every few dispatches the ranges are re-cut from the ranks' own times and the tiles
that change owner take their accumulators along (sharding.migrate_ranges).  Random streams are keyed by the global pixel index and the
frame number and every pixel's samples are added in the same order onto the same running sum, so the image is the same bits whatever
the cuts and whatever N.

torch is plumbing here: device memory for the tile buffer, CUDA events, torch.distributed for NCCL.
"""
import ctypes as C
import hashlib

import numpy as np

from . import sharding


class TiledPathTracer:
    def __init__(self, dev, lc, scenes, width=3840, height=2160, nx=1582, spp_per_dispatch=32, depth=5, block=8, streams=2, rank=0, world=1, dist=None, fast_math=True):
        import torch
        from . import examples_ir
        self.torch, self.dist, self.dev, self.lc = torch, dist, dev, lc
        self.width, self.height, self.rank, self.world = width, height, rank, world
        self.spp_per_dispatch, self.depth = spp_per_dispatch, depth
        self.s = dev.default_stream()
        self.ext = torch.cuda.ExternalStream(self.s.cuda_stream())
        # ---- scene: 10 terrain instances on a 5 x 2 grid (yaw 36 deg * k) + one emissive quad above ----
        verts, tris = scenes.terrain(nx)
        quad_v = np.array([[0, 0, 0], [1, 0, 0], [1, 0, 1], [0, 0, 1]], np.float32)
        quad_t = np.array([[0, 1, 2], [0, 2, 3]], np.uint32)
        self.vb, self.ib = dev.create_buffer_from_array(verts), dev.create_buffer_from_array(tris)
        self.qvb, self.qib = dev.create_buffer_from_array(quad_v), dev.create_buffer_from_array(quad_t)
        self.mesh = dev.create_mesh(self.vb.view(), self.ib.view(), lc.AccelOption())
        self.quad = dev.create_mesh(self.qvb.view(), self.qib.view(), lc.AccelOption())
        self.mesh.build(lc.AccelBuildRequest.FORCE_BUILD); self.quad.build(lc.AccelBuildRequest.FORCE_BUILD)
        self.mesh.build(lc.AccelBuildRequest.FORCE_BUILD)
        self.blas_ms = self.mesh.stats()["build_ms"]
        self.triangles = int(tris.shape[0]) * 10 + 2
        self.accel = dev.create_accel(lc.AccelOption())
        for k in range(10):
            t = np.eye(4, dtype=np.float32); t[:3, :] = scenes.rotation_y(36.0 * k)
            t[:3, 3] = [1.2 * (k % 5), 0.0, 1.2 * (k // 5)]
            self.accel.push_mesh(self.mesh, t)
        l_pos, l_u, l_v = (2.0, 1.6, 0.2), (2.0, 0.0, 0.0), (0.0, 0.0, 1.6)
        t = np.eye(4, dtype=np.float32); t[0, 0], t[2, 2] = l_u[0], l_v[2]; t[:3, 3] = l_pos
        self.accel.push_mesh(self.quad, t)
        self.accel.build(lc.AccelBuildRequest.FORCE_BUILD)
        self.tlas_ms = self.accel.stats()["build_ms"]
        n_inst = 11
        self.vheap, self.iheap = dev.create_bindless_array(n_inst), dev.create_bindless_array(n_inst)
        for i in range(n_inst):
            self.vheap.emplace_buffer_async(i, self.vb if i < 10 else self.qvb); self.iheap.emplace_buffer_async(i, self.ib if i < 10 else self.qib)
        self.s.submit([self.vheap.update_async(), self.iheap.update_async()])
        cam_o, cam_at = np.float32([3.0, 2.5, -3.0]), np.float32([3.0, 0.0, 1.0])
        f = cam_at - cam_o; f /= np.linalg.norm(f)
        r = np.cross(f, np.float32([0, 1, 0])); r /= np.linalg.norm(r)
        u = np.cross(r, f)
        camera = (tuple(map(float, cam_o)), tuple(map(float, f)), tuple(map(float, r)), tuple(map(float, u)), float(np.tan(np.radians(45.0) / 2)))
        light = (l_pos, l_u, l_v, (60.0, 54.0, 45.0), 10)
        self.kernel = examples_ir.tiled_path_tracer_kernel(self.vheap.handle.id, self.iheap.handle.id, camera, light, n_inst, spp_per_dispatch, depth, block=block)
        # the same kernel with rays counted per tile, for the cost pass only (one more atomic per thread costs 10 % of a frame: profiles/r02u_c5_n2.jsonl)
        self.count_kernel = examples_ir.tiled_path_tracer_kernel(self.vheap.handle.id, self.iheap.handle.id, camera, light, n_inst, spp_per_dispatch, depth, block=block, tile_counters=True) if world > 1 else None
        # enable_fast_math is the frontend's default (KernelBuildOptions, runtime/kernel.rs:556-570); hits do not depend on it (csrc/shader.cu)
        self.shader = dev.create_shader(C.addressof(self.kernel.km), fast_math=fast_math, keep=self.kernel)
        self.count_shader = dev.create_shader(C.addressof(self.count_kernel.km), fast_math=fast_math, keep=self.count_kernel) if world > 1 else None
        # ---- tiles ----
        self.tile = sharding.TILE
        self.order_tx, self.order_ty = sharding.tile_order(width, height)
        self.n_tiles = int(self.order_tx.shape[0])
        self.tiles_x = (width + self.tile - 1) // self.tile
        self.all_tile_ids = dev.create_buffer_from_array((self.order_ty * self.tiles_x + self.order_tx).astype(np.uint32))   # Morton order
        self.cost = np.ones(self.n_tiles)
        self.lanes = [self.s] + [dev.create_stream() for _ in range(max(1, streams) - 1)]
        self.joins = [dev.create_event() for _ in self.lanes[1:]]   # one timeline per side lane (a timeline event is a max counter)
        self.serial = 0
        self.tile_ids = (self.order_ty * self.tiles_x + self.order_tx).astype(np.int64)   # global tile id of every Morton position
        self.counters_t = torch.zeros(2 + self.n_tiles, dtype=torch.int64, device="cuda")   # [0] closest-hit, [1] any-hit rays, [2 + tile] rays of a tile
        self.counters = dev.wrap_device_memory(self.counters_t.data_ptr(), 2 + self.n_tiles, 8, 8)
        self.out_t = self.out = None
        self.set_bounds(sharding.balanced_bounds(self.cost, world))

    def set_bounds(self, bounds, clear=True):
        """this rank renders tiles [bounds[rank], bounds[rank + 1]) of the Morton order.  The tile buffer holds the WHOLE frame in Morton
        order (133 MB at 4K; a rank is authoritative for its own range only) plus one longest-possible range of padding, so that the
        equal-count all-gather can always read `cap` tiles from the start of any range."""
        torch = self.torch
        self.bounds = np.asarray(bounds, np.int64)
        self.b0, self.b1 = int(self.bounds[self.rank]), int(self.bounds[self.rank + 1])
        if self.out_t is None:
            cap = 2 * self.n_tiles * self.tile * self.tile
            self.out_t = torch.zeros((cap, 4), dtype=torch.float32, device="cuda")
            self.out = self.dev.wrap_device_memory(self.out_t.data_ptr(), cap, 16, 16)
        elif clear:
            self.out_t.zero_()
        torch.cuda.synchronize()

    def render(self, n_dispatch, first_frame, shader=None):
        """enqueue n_dispatch passes of spp_per_dispatch samples over this rank's range; the range is split over the lanes (streams) so
        that the tail of one dispatch overlaps the other lane's work; all lanes are joined into the default stream"""
        tile, n = self.tile, self.b1 - self.b0
        edges = np.linspace(0, n, len(self.lanes) + 1).astype(int)
        for li, lane in enumerate(self.lanes):
            a, b = int(edges[li]), int(edges[li + 1])
            if b == a:
                continue
            lane.submit([(shader or self.shader).dispatch_async((tile, tile * (b - a)), self.all_tile_ids.view(self.b0 + a, b - a), self.out.view((self.b0 + a) * tile * tile, (b - a) * tile * tile), self.accel,
                                                    np.array([self.width, self.height, first_frame + i, b - a], np.uint32), self.counters) for i in range(n_dispatch)])
        self.serial += 1
        for lane, join in zip(self.lanes[1:], self.joins):
            join.signal(lane, self.serial); join.wait(self.s, self.serial)

    def timed(self, fn):
        torch = self.torch
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(self.ext); fn(); e1.record(self.ext); self.s.synchronize()
        return e0.elapsed_time(e1)

    def all_times(self, ms):
        torch = self.torch
        if self.world == 1:
            return [ms]
        t = torch.tensor([ms], device="cuda"); out = torch.empty(self.world, device="cuda")
        self.dist.all_gather_into_tensor(out, t)
        return [float(x) for x in out.tolist()]

    def probe_cost(self, pieces_per_rank=8, frame0=90000):
        """first estimate of the cost map: the Morton curve is cut into world * pieces_per_rank equal pieces dealt round-robin, every
        rank times each of its pieces on its own (one small dispatch each, samples discarded) and the piece costs are exchanged — a
        piecewise-constant map with world * pieces_per_rank steps, which already sees that sky tiles cost a fifth of terrain tiles"""
        torch = self.torch
        n_pieces = self.world * pieces_per_rank
        edges = np.linspace(0, self.n_tiles, n_pieces + 1).astype(np.int64)
        saved = (self.bounds, self.b0, self.b1)
        mine = []
        for j in range(pieces_per_rank):
            piece = j * self.world + self.rank
            self.b0, self.b1 = int(edges[piece]), int(edges[piece + 1])
            mine.append(self.timed(lambda: self.render(1, frame0 + j)) if self.b1 > self.b0 else 0.0)
        self.bounds, self.b0, self.b1 = saved
        t = torch.tensor(mine, device="cuda", dtype=torch.float32)
        if self.world > 1:
            allt = torch.empty(self.world * pieces_per_rank, device="cuda", dtype=torch.float32)
            self.dist.all_gather_into_tensor(allt, t)
            allt = allt.view(self.world, pieces_per_rank).cpu().numpy()
        else:
            allt = t.view(1, -1).cpu().numpy()
        for piece in range(n_pieces):
            r, j = piece % self.world, piece // self.world
            b0, b1 = int(edges[piece]), int(edges[piece + 1])
            if b1 > b0:
                self.cost[b0:b1] = max(float(allt[r, j]), 1e-6) / (b1 - b0)
        self.set_bounds(sharding.balanced_bounds(self.cost, self.world))
        return True

    def cost_from_ray_counts(self, dispatches=1, frame0=95000):
        """the cost map at tile resolution from ONE short pass: every rank renders its current range, the kernel counts the rays it traced
        per tile, the counts are summed over the ranks (each tile has one owner) and scaled, range by range, to the time that range took —
        rays are what a tile costs, up to how expensive a ray is where the range looks"""
        torch = self.torch
        self.counters_t.zero_()
        ms = self.timed(lambda: self.render(dispatches, frame0, self.count_shader))
        times = self.all_times(ms)
        counts = self.counters_t[2:].clone()
        if self.world > 1:
            self.dist.all_reduce(counts, op=self.dist.ReduceOp.SUM)
        per_tile = counts.cpu().numpy().astype(np.float64)[self.tile_ids]   # Morton order
        self.cost = sharding.refine_cost(np.maximum(per_tile, 1.0), self.bounds, times)
        self.recut_to(sharding.balanced_bounds(self.cost, self.world))
        self.counters_t.zero_(); self.out_t.zero_()
        return times

    def balance(self, passes=4, dispatches=2, frame0=100000):
        """refine the cost map from per-rank times of short passes (`dispatches` dispatches each, samples discarded) and re-cut the
        ranges; returns the history [(imbalance = max / mean of the per-rank times, tiles per rank, ms per rank)]"""
        history = []
        for p in range(passes):
            ms = self.timed(lambda: self.render(dispatches, frame0 + p * dispatches))
            times = self.all_times(ms)
            history.append((max(times) / (sum(times) / len(times)), [int(x) for x in np.diff(self.bounds)], [round(t, 2) for t in times]))
            self.recut(times)   # the same path the frame takes between its portions (also warms up the point-to-point connections)
        self.out_t.zero_()
        return history

    def gather(self):
        """the one collective of the result path: all-gather of every rank's range, padded to the longest range (equal counts), read from
        the rank's full-frame buffer (stream-ordered after the render on the default stream)"""
        torch = self.torch
        T = self.tile * self.tile
        cap = int(np.max(np.diff(self.bounds))) * T
        local = self.out_t[self.b0 * T: self.b0 * T + cap]
        if self.world == 1:
            return local[None]
        if getattr(self, "_gathered", None) is None or self._gathered.shape[1] != cap:
            self._gathered = torch.empty((self.world, cap, 4), dtype=self.out_t.dtype, device="cuda")
        with torch.cuda.stream(self.ext):
            self.dist.all_gather_into_tensor(self._gathered.view(-1), local.reshape(-1))
        return self._gathered

    def recut(self, times, threshold=1.0):
        """refine the cost map from the ranks' times for the current ranges and, when they differ by more than `threshold` (max / mean),
        re-cut and move the accumulators of the tiles that change owner (in-frame re-balancing); returns the tiles this rank sent or received"""
        self.cost = sharding.refine_cost(self.cost, self.bounds, times)
        if max(times) / (sum(times) / len(times)) <= threshold:
            return 0
        return self.recut_to(sharding.balanced_bounds(self.cost, self.world))

    def recut_to(self, new_bounds):
        if np.array_equal(new_bounds, self.bounds):
            return 0
        with self.torch.cuda.stream(self.ext):
            moved = sharding.migrate_ranges(self.out_t, self.bounds, new_bounds, self.rank, self.dist, self.tile * self.tile)
        self.bounds = np.asarray(new_bounds, np.int64)
        self.b0, self.b1 = int(self.bounds[self.rank]), int(self.bounds[self.rank + 1])
        return moved

    def frame(self, spp, first_frame=0, recuts=0):
        """one frame: render spp samples per pixel AND gather the framebuffer, timed as one region on the device; returns (ms, gathered,
        dispatches).  recuts > 0 (world > 1): the dispatches are issued in recuts + 1 portions and the ranges are re-cut between them
        from each portion's per-rank device time (recut()); the exchanges and migrations are inside the timed region."""
        torch = self.torch
        n_dispatch = max(1, spp // self.spp_per_dispatch)
        self.out_t.zero_(); self.counters_t.zero_(); torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        parts = 1 + (recuts if self.world > 1 else 0)
        edges = np.linspace(0, n_dispatch, min(parts, n_dispatch) + 1).astype(int)
        self.recut_log = []
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(self.ext)
        for k in range(len(edges) - 1):
            a, b = int(edges[k]), int(edges[k + 1])
            if k + 2 == len(edges):          # last portion: nothing left to re-balance for
                self.render(b - a, first_frame + a)
            else:
                ms = self.timed(lambda: self.render(b - a, first_frame + a))
                times = self.all_times(ms)
                moved = self.recut(times, threshold=1.02)
                self.recut_log.append((round(max(times) / (sum(times) / len(times)), 3), moved))
        if self.world > 1 and len(self.lanes) > 2:
            # belt and braces for the unverified combination (known issue: one 8-rank run with 4 lanes gathered tiles whose
            # lanes had not finished although every lane is joined into the default stream by a timeline event): the host waits for the
            # side lanes before it enqueues the gather
            for lane in self.lanes[1:]:
                lane.synchronize()
        g = self.gather()
        e1.record(self.ext); self.s.synchronize()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1), g, n_dispatch

    def image(self, gathered):
        return sharding.untile_ranges(gathered.cpu().numpy(), self.width, self.height, self.bounds)

    @staticmethod
    def sha(img):
        return hashlib.sha256(np.ascontiguousarray(img).tobytes()).hexdigest()

    def destroy(self):
        for r in ((self.shader, self.count_shader) if self.count_shader else (self.shader,)) + (self.vheap, self.iheap, self.all_tile_ids, self.out, self.counters, self.accel, self.mesh, self.quad, self.vb, self.ib, self.qvb, self.qib):
            r.destroy()
        for lane in self.lanes[1:]:
            lane.destroy()
        for j in self.joins:
            j.destroy()
