"""CPU tests of the tile scheduler (host logic of the multi-GPU partition), including a
world_size-2 gloo run: the parts of the grid are disjoint, cover it, and balance."""
import os
import socket

import numpy as np
import pytest

import tomahawk_b200 as tb
from tomahawk_b200 import synth


def _meta(n, step=100):
    m = np.zeros(n, dtype=tb.VARIANT_DTYPE)
    m["pos"] = np.arange(n) * step
    m["ac"] = 10
    m["hwe"] = 1.0
    return m


def _pairs_of(tiles, ti, tj, n, diag=True):
    seen = set()
    for i0, j0 in tiles.tolist():
        for i in range(i0, min(i0 + ti, n)):
            for j in range(j0, min(j0 + tj, n)):
                if not diag or i < j:
                    seen.add((i, j))
    return seen


@pytest.mark.parametrize("n,ti,tj", [(300, 128, 128), (1000, 64, 128), (257, 64, 64), (5, 128, 128)])
def test_all_pairs_cover_exactly_once(n, ti, tj):
    s = tb.default_settings(force_phased=1)
    tiles, pairs = tb.plan_tiles(s, _meta(n), ti, tj)
    assert pairs == n * (n - 1) // 2
    assert len(set(map(tuple, tiles.tolist()))) == len(tiles)
    if n <= 300:
        assert _pairs_of(tiles, ti, tj, n) == {(i, j) for i in range(n) for j in range(i + 1, n)}


@pytest.mark.parametrize("parts", [2, 3, 8])
def test_parts_are_disjoint_and_cover(parts):
    n, ti, tj = 3000, 128, 128
    meta = _meta(n)
    whole, pairs_whole = tb.plan_tiles(tb.default_settings(), meta, ti, tj)
    got, pairs = [], []
    for r in range(parts):
        t, p = tb.plan_tiles(tb.default_settings(part_index=r, part_count=parts), meta, ti, tj)
        got.append(set(map(tuple, t.tolist())))
        pairs.append(p)
    assert sum(pairs) == pairs_whole == n * (n - 1) // 2
    union = set().union(*got)
    assert union == set(map(tuple, whole.tolist()))
    assert sum(len(g) for g in got) == len(union)
    assert max(pairs) <= 1.5 * min(pairs)


def test_chunks_partition_the_triangle():
    # -c 3 = 2x2 block grid: (0,0) diag, (0,1) square, (1,1) diag; reference ld_balancing.h:45-79
    n = 2000
    meta = _meta(n)
    total = 0
    seen = set()
    for c in range(3):
        s = tb.default_settings(n_chunks=3, c_chunk=c)
        tiles, pairs = tb.plan_tiles(s, meta, 128, 128)
        total += pairs
    assert total == n * (n - 1) // 2


def test_window_band_prunes_far_tiles():
    n = 20000
    meta = _meta(n, step=100)
    s_all = tb.default_settings()
    s_win = tb.default_settings(window=1, l_window=50000, force_phased=1)
    all_tiles, _ = tb.plan_tiles(s_all, meta, 128, 128)
    win_tiles, _ = tb.plan_tiles(s_win, meta, 128, 128)
    assert 0 < len(win_tiles) < 0.1 * len(all_tiles)
    # auto mode (neither -p nor -u): the reference has no per-pair window test, only the balancer's row prune
    # over 500-variant blocks (ld_balancing.h:189-196) -- a coarser band that contains the -p one
    auto_tiles, _ = tb.plan_tiles(tb.default_settings(window=1, l_window=50000), meta, 128, 128)
    assert len(win_tiles) < len(auto_tiles) < 0.2 * len(all_tiles)
    assert set(map(tuple, win_tiles.tolist())) <= set(map(tuple, auto_tiles.tolist()))
    # every in-window pair's tile is kept
    keep = set(map(tuple, win_tiles.tolist()))
    for i in (0, 777, 10000, 19000):
        for j in (i + 1, i + 200, min(i + 499, n - 1)):
            if j < n and j > i:
                assert ((i // 128) * 128, (j // 128) * 128) in keep


def _worker(rank, world, port, q):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import torch

    n = 2500
    # rank 0 owns the matrix metadata and broadcasts it (the NCCL broadcast of bench.py, on gloo)
    meta = _meta(n) if rank == 0 else np.zeros(n, dtype=tb.VARIANT_DTYPE)
    t = torch.from_numpy(meta.view(np.uint8).copy())
    dist.broadcast(t, src=0)
    meta = t.numpy().view(tb.VARIANT_DTYPE)
    tiles, pairs = tb.plan_tiles(tb.default_settings(part_index=rank, part_count=world), meta, 128, 128)
    gathered = [None] * world
    dist.all_gather_object(gathered, (tiles.tolist(), pairs))
    if rank == 0:
        q.put(gathered)
    dist.destroy_process_group()


def test_two_rank_gloo_partition():
    import torch.multiprocessing as mp

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    gathered = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (t0, p0), (t1, p1) = gathered
    n = 2500
    assert p0 + p1 == n * (n - 1) // 2
    a, b = set(map(tuple, t0)), set(map(tuple, t1))
    assert not (a & b)
    whole, _ = tb.plan_tiles(tb.default_settings(), _meta(n), 128, 128)
    assert a | b == set(map(tuple, whole.tolist()))


@pytest.mark.parametrize("parts", [2, 8])
def test_large_grid_parts_take_whole_super_tiles(parts):
    """At bench scale (C2, 256 x 240 tiles) a part owns whole 32 x 32 super-tiles: disjoint cover,
    tile counts balanced to within 2 %, and each part touches far fewer operand row groups per tile
    than round-robin dealing of single tiles would (the L2 / DRAM locality the dealing exists for)."""
    n, ti, tj = 200_000, 256, 240
    meta = _meta(n)
    whole, pairs_whole = tb.plan_tiles(tb.default_settings(), meta, ti, tj)
    key = lambda t: (t[:, 0].astype(np.uint64) << np.uint64(32)) | t[:, 1].astype(np.uint64)
    all_keys, counts, pairs = [], [], []
    for r in range(parts):
        t, p = tb.plan_tiles(tb.default_settings(part_index=r, part_count=parts), meta, ti, tj)
        all_keys.append(key(t))
        counts.append(len(t))
        pairs.append(p)
        # locality: tiles per distinct (row group, column group) the part has to stream
        groups = len(np.unique(t[:, 0])) + len(np.unique(t[:, 1]))
        supers = len(np.unique(key(np.stack([t[:, 0] // (32 * ti), t[:, 1] // (32 * tj)], axis=1))))
        assert len(t) / supers > 400, "a part's tiles should fill its super-tiles"
        assert groups <= whole[:, 0].max() // ti + whole[:, 1].max() // tj + 2
    cat = np.concatenate(all_keys)
    assert len(np.unique(cat)) == len(cat) == len(whole)
    assert np.array_equal(np.sort(cat), np.sort(key(whole)))
    assert sum(pairs) == pairs_whole == n * (n - 1) // 2
    assert max(counts) <= 1.02 * min(counts)
