"""Host-side logic of the multi-GPU partitioning, on CPU: ray slices, Morton tile assignment and the one
collective of the path (all-gather of equal-sized per-rank tile buffers) over gloo with world_size 2."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import luisa_compute_rs_b200.sharding as sh


def test_ray_slices_partition_the_batch():
    for n in (0, 1, 7, 1000, (1 << 24) + 3):
        for world in (1, 2, 3, 4, 8):
            edges = [sh.ray_slice(n, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(edges[r][1] == edges[r + 1][0] for r in range(world - 1))
            sizes = [e - b for b, e in edges]
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.parametrize("w,h", [(3840, 2160), (1024, 1024), (100, 70), (64, 64)])
@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_tiles_cover_every_pixel_exactly_once(w, h, world):
    seen = np.zeros(w * h, np.int32)
    counts = []
    for r in range(world):
        tx, ty = sh.tiles_of_rank(w, h, r, world)
        idx, valid = sh.pixels_of_tiles(tx, ty, w, h)
        np.add.at(seen, idx[valid], 1)
        counts.append(tx.shape[0])
        assert tx.shape[0] <= sh.padded_tile_count(w, h, world)
    assert (seen == 1).all()
    assert max(counts) - min(counts) <= 1   # balanced along the Morton curve


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, w, h, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # every rank "renders" its tiles: pixel value = f(global pixel index), independent of the partition
    tx, ty = sh.tiles_of_rank(w, h, rank, world)
    idx, valid = sh.pixels_of_tiles(tx, ty, w, h)
    per_rank = sh.padded_tile_count(w, h, world) * sh.TILE * sh.TILE
    local = torch.zeros(per_rank, 4)
    vals = torch.from_numpy(np.stack([idx, idx * 2, idx % 7, np.ones_like(idx)], 1).astype(np.float32))
    vals[~torch.from_numpy(valid)] = 0
    local[: idx.shape[0]] = vals
    g = sh.gather_tiles(local, dist, world)
    img = sh.untile(g.numpy(), w, h, world)
    np.save(os.path.join(out_dir, f"img{rank}.npy"), img)
    # ray batches: each rank contributes its contiguous slice of a global hit buffer
    n = 1001
    b, e = sh.ray_slice(n, rank, world)
    pad = -(-n // world)
    mine = torch.full((pad,), -1, dtype=torch.int64)
    mine[: e - b] = torch.arange(b, e)
    allr = sh.gather_tiles(mine, dist, world)
    flat = torch.cat([allr[r, : sh.ray_slice(n, r, world)[1] - sh.ray_slice(n, r, world)[0]] for r in range(world)])
    assert torch.equal(flat, torch.arange(n))
    dist.destroy_process_group()


def test_gloo_world2_gather_reassembles_the_framebuffer(tmp_path):
    w, h, world = 200, 136, 2
    mp.spawn(_worker, args=(world, _free_port(), w, h, str(tmp_path)), nprocs=world, join=True)
    px = np.arange(w * h)
    want = np.stack([px, px * 2, px % 7, np.ones_like(px)], 1).astype(np.float32).reshape(h, w, 4)
    for r in range(world):
        assert np.array_equal(np.load(tmp_path / f"img{r}.npy"), want)


# ---- cost-balanced contiguous partition (sharding.balanced_bounds / refine_cost / untile_ranges) ---------------------------------------
def test_balanced_bounds_cut_equal_cost_and_keep_ranges_non_empty():
    rng = np.random.default_rng(0)
    cost = np.exp(rng.normal(size=2040) * 0.5) * (1 + 2 * np.sin(np.arange(2040) / 300) ** 2)
    for world in (1, 2, 3, 4, 8):
        b = sh.balanced_bounds(cost, world)
        assert b[0] == 0 and b[-1] == 2040 and np.all(np.diff(b) > 0) and b.shape[0] == world + 1
        sums = np.array([cost[b[r]:b[r + 1]].sum() for r in range(world)])
        assert sums.max() / sums.mean() < 1.01            # one tile is ~0.05 % of the total
    # degenerate inputs: all cost in one tile, fewer tiles than ranks is not asked for, exactly `world` tiles
    spike = np.full(16, 1e-9); spike[5] = 1.0
    b = sh.balanced_bounds(spike, 4)
    assert np.all(np.diff(b) > 0) and b[-1] == 16
    assert np.array_equal(sh.balanced_bounds(np.ones(8), 8), np.arange(9))


def test_cost_map_refinement_converges_from_per_rank_times_only():
    """what the renderer does between passes: only ONE number per rank is measured, yet the cut converges to balance"""
    rng = np.random.default_rng(1)
    true = np.exp(rng.normal(size=2040) * 0.5) * (1 + 3 * (np.arange(2040) > 1400))
    for world in (2, 4, 8):
        cost = np.ones(2040)
        imbalance = []
        for _ in range(5):
            b = sh.balanced_bounds(cost, world)
            times = [true[b[r]:b[r + 1]].sum() for r in range(world)]
            imbalance.append(max(times) / np.mean(times))
            cost = sh.refine_cost(cost, b, times)
        assert imbalance[0] > 1.2 and imbalance[-1] < 1.03, imbalance


def _range_worker(rank, world, port, w, h, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n_tiles = sh.tile_order(w, h)[0].shape[0]
    cost = 1.0 + np.arange(n_tiles) % 5                      # every rank holds the same map (it is built from all-gathered times)
    bounds = sh.balanced_bounds(cost, world)
    tx, ty = sh.tiles_of_range(w, h, int(bounds[rank]), int(bounds[rank + 1]))
    idx, valid = sh.pixels_of_tiles(tx, ty, w, h)
    per_rank = int(np.max(np.diff(bounds))) * sh.TILE * sh.TILE   # ranks pad to the longest range: one all-gather of equal counts
    local = torch.zeros(per_rank, 4)
    vals = torch.from_numpy(np.stack([idx, idx * 2, idx % 7, np.ones_like(idx)], 1).astype(np.float32))
    vals[~torch.from_numpy(valid)] = 0
    local[: idx.shape[0]] = vals
    g = sh.gather_tiles(local, dist, world)
    np.save(os.path.join(out_dir, f"img{rank}.npy"), sh.untile_ranges(g.numpy(), w, h, bounds))
    # the per-rank times travel the same way: one value per rank
    t = sh.gather_tiles(torch.tensor([float(rank + 1)]), dist, world)
    assert t.reshape(-1).tolist() == [float(r + 1) for r in range(world)]
    dist.destroy_process_group()


def test_gloo_world2_gather_of_unequal_contiguous_ranges(tmp_path):
    w, h, world = 200, 136, 2
    mp.spawn(_range_worker, args=(world, _free_port(), w, h, str(tmp_path)), nprocs=world, join=True)
    px = np.arange(w * h)
    want = np.stack([px, px * 2, px % 7, np.ones_like(px)], 1).astype(np.float32).reshape(h, w, 4)
    for r in range(world):
        assert np.array_equal(np.load(tmp_path / f"img{r}.npy"), want)


def test_migration_plan_moves_every_reassigned_tile_exactly_once():
    rng = np.random.default_rng(3)
    for world in (2, 4, 8):
        for _ in range(20):
            n = 500
            old = np.concatenate([[0], np.sort(rng.choice(np.arange(1, n), world - 1, replace=False)), [n]])
            new = np.concatenate([[0], np.sort(rng.choice(np.arange(1, n), world - 1, replace=False)), [n]])
            owner_old = np.searchsorted(old, np.arange(n), side="right") - 1
            owner_new = np.searchsorted(new, np.arange(n), side="right") - 1
            moved = np.zeros(n, int)
            for src, dst, a, b in sh.migration_plan(old, new):
                assert src != dst and np.all(owner_old[a:b] == src) and np.all(owner_new[a:b] == dst)
                moved[a:b] += 1
            assert np.array_equal(moved, (owner_old != owner_new).astype(int))


def _migrate_worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n_tiles, per = 40, 6
    old, new = np.array([0, 25, 40]), np.array([0, 12, 40])
    buf = torch.full((n_tiles * per,), -1.0)                      # a rank is authoritative for its own range only
    buf[old[rank] * per: old[rank + 1] * per] = torch.arange(old[rank] * per, old[rank + 1] * per, dtype=torch.float32) + 1000 * rank
    moved = sh.migrate_ranges(buf, old, new, rank, dist, per)
    assert moved == 13
    mine = buf[new[rank] * per: new[rank + 1] * per]
    want = torch.arange(new[rank] * per, new[rank + 1] * per, dtype=torch.float32) + 1000 * torch.from_numpy((np.arange(new[rank] * per, new[rank + 1] * per) // per >= 25).astype(np.float32))
    np.save(os.path.join(out_dir, f"ok{rank}.npy"), np.array([bool(torch.equal(mine, want))]))
    dist.destroy_process_group()


def test_gloo_world2_accumulators_follow_their_tiles_when_the_ranges_are_recut(tmp_path):
    mp.spawn(_migrate_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    assert all(bool(np.load(tmp_path / f"ok{r}.npy")[0]) for r in range(2))
