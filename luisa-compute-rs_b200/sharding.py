"""Multi-GPU partitioning of the ray-query path: one process per GPU, meshes + Accel replicated
(every rank runs the deterministic build itself from the same vertex/index data — no broadcast), rays or image
tiles partitioned, and ONE collective: the gather of per-rank results (framebuffer tiles / hit records) over NCCL.

Host-side logic only; torch.distributed is the plumbing (gloo in CPU tests, NCCL on the GPUs).
"""
import numpy as np

TILE = 64  # pixels per tile edge


def ray_slice(n_rays, rank, world):
    """Contiguous, balanced [begin, end) slice of a ray batch for `rank`."""
    base, rem = divmod(n_rays, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def _part1by1(v):
    v = v.astype(np.uint64) & np.uint64(0xFFFF)
    v = (v | (v << np.uint64(8))) & np.uint64(0x00FF00FF)
    v = (v | (v << np.uint64(4))) & np.uint64(0x0F0F0F0F)
    v = (v | (v << np.uint64(2))) & np.uint64(0x33333333)
    v = (v | (v << np.uint64(1))) & np.uint64(0x55555555)
    return v


def tile_order(width, height, tile=TILE):
    """Tiles of the image sorted along a Morton curve; returns (tx, ty) arrays in assignment order."""
    nx, ny = (width + tile - 1) // tile, (height + tile - 1) // tile
    ty, tx = np.divmod(np.arange(nx * ny), nx)
    code = _part1by1(tx) | (_part1by1(ty) << np.uint64(1))
    order = np.argsort(code, kind="stable")
    return tx[order], ty[order]


def tiles_of_rank(width, height, rank, world, tile=TILE, chunk=1):
    """Round-robin along the Morton curve in runs of `chunk` tiles: tile k of the curve belongs to rank (k // chunk) % world.
    chunk = 1 balances best; larger runs keep a rank's rays in one part of the scene (better L2 reuse of the BVH)."""
    tx, ty = tile_order(width, height, tile)
    sel = (np.arange(tx.shape[0]) // chunk) % world == rank
    return tx[sel], ty[sel]


def pixels_of_tiles(tx, ty, width, height, tile=TILE):
    """Global pixel indices (row-major) of the given tiles, tile after tile, clipped at the image border, plus a
    validity mask for the padded tail (every tile contributes tile*tile entries so ranks have equal counts)."""
    oy, ox = np.divmod(np.arange(tile * tile), tile)
    px = tx[:, None] * tile + ox[None, :]
    py = ty[:, None] * tile + oy[None, :]
    valid = (px < width) & (py < height)
    idx = np.where(valid, py * width + px, 0)
    return idx.reshape(-1), valid.reshape(-1)


def padded_tile_count(width, height, world, tile=TILE, chunk=1):
    """Largest number of tiles any rank owns (ranks pad their buffers to it so that the all-gather has equal counts)."""
    nx, ny = (width + tile - 1) // tile, (height + tile - 1) // tile
    if chunk == 1:
        return -(-(nx * ny) // world)
    return max(int(((np.arange(nx * ny) // chunk) % world == r).sum()) for r in range(world))


def gather_tiles(local, dist_module, world):
    """all_gather of equally sized per-rank tile buffers (torch tensors).  Returns a (world, ...) tensor on every rank."""
    import torch
    out = torch.empty((world,) + tuple(local.shape), dtype=local.dtype, device=local.device)
    dist_module.all_gather_into_tensor(out.view(-1), local.contiguous().view(-1))
    return out


def untile(gathered, width, height, world, tile=TILE, chunk=1):
    """Scatter gathered per-rank tile buffers (numpy, shape (world, tiles_per_rank*tile*tile, C)) back into an image."""
    channels = gathered.shape[-1]
    img = np.zeros((height * width, channels), dtype=gathered.dtype)
    per_rank = padded_tile_count(width, height, world, tile, chunk)
    for r in range(world):
        tx, ty = tiles_of_rank(width, height, r, world, tile, chunk)
        idx, valid = pixels_of_tiles(tx, ty, width, height, tile)
        buf = gathered[r, : tx.shape[0] * tile * tile]
        img[idx[valid]] = buf[valid]
        assert tx.shape[0] <= per_rank
    return img.reshape(height, width, channels)


# ---- cost-balanced contiguous partition of the Morton curve -------------------------------------------------------------------------
# Round-robin single tiles balance a path tracer's load but scatter every rank's rays over the whole scene (a rank re-reads the whole
# BVH through its L2); contiguous runs keep a rank in one part of the scene but are unbalanced, because cost per tile varies with what
# the tile sees.  The partition below has both: rank r renders the contiguous range [bounds[r], bounds[r + 1]) of the Morton-ordered
# tiles, and the bounds are cut so that every range carries the same estimated cost.  The estimate is refined from what a progressive
# renderer already has — the time each rank took for its range in the previous pass: the per-tile estimate is rescaled inside every
# range so that the range sums to the measured time, then the bounds are re-cut.  Pixels never depend on the partition (random streams
# are keyed by the global pixel index), so the image is the same bits for every cut.

def balanced_bounds(cost, world):
    """cost: per-tile estimates along the Morton curve (positive).  Returns world + 1 tile indices; range r = [b[r], b[r + 1]).  Every
    range is non-empty when there are at least `world` tiles."""
    cost = np.maximum(np.asarray(cost, np.float64), 1e-30)
    n = cost.shape[0]
    cum = np.concatenate([[0.0], np.cumsum(cost)])
    targets = cum[-1] * np.arange(1, world) / world
    inner = np.searchsorted(cum, targets, side="left")
    # the cut goes to whichever side of the crossing tile leaves the smaller error
    inner = np.where((inner > 0) & (np.abs(cum[np.maximum(inner - 1, 0)] - targets) < np.abs(cum[np.minimum(inner, n)] - targets)), inner - 1, inner)
    bounds = np.concatenate([[0], inner, [n]]).astype(np.int64)
    if n >= world:   # keep every range non-empty
        for r in range(1, world):
            bounds[r] = max(bounds[r], bounds[r - 1] + 1)
        for r in range(world - 1, 0, -1):
            bounds[r] = min(bounds[r], bounds[r + 1] - 1)
    return bounds


def refine_cost(cost, bounds, times):
    """Rescale the per-tile estimates inside every range so that the range sums to the time measured for it."""
    cost = np.array(cost, np.float64)
    for r, t in enumerate(times):
        b0, b1 = int(bounds[r]), int(bounds[r + 1])
        if b1 > b0 and t > 0:
            total = cost[b0:b1].sum()
            cost[b0:b1] = cost[b0:b1] * (t / total) if total > 0 else t / (b1 - b0)
    return cost


def tiles_of_range(width, height, begin, end, tile=TILE):
    tx, ty = tile_order(width, height, tile)
    return tx[begin:end], ty[begin:end]


def untile_ranges(gathered, width, height, bounds, tile=TILE):
    """untile() for the contiguous partition: gathered[r] holds the tiles [bounds[r], bounds[r + 1]) of the Morton order, padded to the
    longest range."""
    channels = gathered.shape[-1]
    img = np.zeros((height * width, channels), dtype=gathered.dtype)
    tx, ty = tile_order(width, height, tile)
    for r in range(len(bounds) - 1):
        b0, b1 = int(bounds[r]), int(bounds[r + 1])
        idx, valid = pixels_of_tiles(tx[b0:b1], ty[b0:b1], width, height, tile)
        img[idx[valid]] = gathered[r, : (b1 - b0) * tile * tile][valid]
    return img.reshape(height, width, channels)


# ---- re-cutting the ranges while a frame accumulates ---------------------------------------------------------------------------------
# A frame of many dispatches need not live with the cut it started with: every few dispatches the ranks exchange their times, the cost
# map is refined and the ranges are re-cut.  A tile that changes owner takes its accumulator along (one point-to-point copy from the old
# owner to the new one), so every pixel still receives its samples one after the other on top of the same running sum — the image stays
# the same bits.  Every rank keeps a buffer for the WHOLE Morton-ordered frame and is authoritative for its own range only.

def migration_plan(old_bounds, new_bounds):
    """[(src_rank, dst_rank, first_tile, end_tile)]: the tiles of [first, end) belonged to src under old_bounds and belong to dst under
    new_bounds (src != dst).  Every rank computes the same plan."""
    plan = []
    world = len(old_bounds) - 1
    for src in range(world):
        for dst in range(world):
            if src == dst:
                continue
            a = max(int(old_bounds[src]), int(new_bounds[dst])); b = min(int(old_bounds[src + 1]), int(new_bounds[dst + 1]))
            if b > a:
                plan.append((src, dst, a, b))
    return plan


def migrate_ranges(frame_buf, old_bounds, new_bounds, rank, dist_module, elems_per_tile):
    """Move the accumulators of re-assigned tiles between the ranks' full-frame buffers (torch tensor, first dimension = tiles *
    elems_per_tile in Morton order).  One batch of point-to-point transfers; returns the number of tiles this rank sent or received."""
    ops, moved = [], 0
    for src, dst, a, b in migration_plan(old_bounds, new_bounds):
        piece = frame_buf[a * elems_per_tile: b * elems_per_tile]
        if rank == src:
            ops.append(dist_module.P2POp(dist_module.isend, piece, dst)); moved += b - a
        elif rank == dst:
            ops.append(dist_module.P2POp(dist_module.irecv, piece, src)); moved += b - a
    if ops:
        for w in dist_module.batch_isend_irecv(ops):
            w.wait()
    return moved
