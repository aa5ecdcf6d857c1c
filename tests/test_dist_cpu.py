"""N>1 host logic on CPU: world_size-2 gloo processes exercise the image-parallel sharding, the max-over-ranks
timing reduction and the final gather (the data path itself has no collective)."""
import os
import socket

import torch
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_items, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from omgsr_b200 import dist as D
    ws, r, _ = D.init(backend="gloo")
    assert (ws, r) == (world, rank)
    mine = D.shard_indices(n_items, r, ws)
    # "process" each image: a deterministic function of its index stands in for the SR result
    local = {i: torch.full((2, 2), float(i)) for i in mine}
    t = D.max_over_ranks(1.0 + rank)
    total = D.sum_over_ranks(len(mine))
    D.barrier()
    res = D.gather_results(local, n_items)
    if rank == 0:
        out.put((t, total, [float(x[0, 0]) for x in res], mine))
    torch.distributed.destroy_process_group()


def test_image_parallel_sharding_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    n = 7
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    t, total, order, mine0 = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert t == 2.0            # slowest rank
    assert total == n          # every image processed exactly once
    assert order == [float(i) for i in range(n)]
    assert mine0 == [0, 2, 4, 6]


def test_shard_indices_cover_everything_once():
    from omgsr_b200.dist import shard_indices
    for n in (0, 1, 5, 16, 17):
        for w in (1, 2, 4, 8):
            got = sorted(i for r in range(w) for i in shard_indices(n, r, w))
            assert got == list(range(n))


def _worker_gn(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from omgsr_b200 import dist as D
    D.init(backend="gloo")
    g = torch.Generator().manual_seed(7)
    # 5 tiles of one image (n_img = 2 images in the batch): per-tile mean / var / pixel counts
    mean = torch.randn(5, 2, 32, generator=g)
    var = torch.rand(5, 2, 32, generator=g)
    pix = torch.tensor([400.0, 400.0, 380.0, 361.0, 100.0])
    mine, shadow = D.my_tiles(5, rank, world)
    assert not shadow
    num = torch.stack([sum(pix[t] * mean[t] for t in mine), sum(pix[t] * var[t] for t in mine)], dim=-1)  # [2,32,2]
    wsum = torch.full((2,), float(sum(pix[t] for t in mine)))
    shared = D.reduce_gn_partials(num, wsum)
    ref_mean = (pix[:, None, None] * mean).sum(0) / pix.sum()
    ref_var = (pix[:, None, None] * var).sum(0) / pix.sum()
    ok = torch.allclose(shared[..., 0], ref_mean, atol=1e-5) and torch.allclose(shared[..., 1], ref_var, atol=1e-5)
    # tile-parallel batch forward: order restored after the all-gather, padding entries dropped
    D.enable_tile_parallel(True)
    batch = torch.arange(9, dtype=torch.float32).reshape(9, 1) * 10
    got = D.tile_parallel_forward(lambda t: t + 1, batch)
    ok = ok and torch.equal(got, batch + 1)
    # 9 tiles on 2 ranks with a cooperative form for the left-over tile (the sequence-parallel FLUX forward stands
    # behind fn_all on the GPU): 8 tiles in 4 full waves, the ninth computed by both ranks together
    calls = {"fn": 0, "all": 0}

    def fn(t):
        calls["fn"] += t.shape[0]
        return t * 2

    def fn_all(t):  # each rank handles half of the row, one all-gather puts it together (what the engine does with K|V)
        calls["all"] += 1
        half = t.shape[1] // world
        part = (t[:, rank * half:(rank + 1) * half] * 2).contiguous()
        return D.all_gather_batch(part).permute(1, 0, 2).reshape(t.shape[0], -1)

    batch2 = torch.arange(9 * 4, dtype=torch.float32).reshape(9, 4)
    got2 = D.tile_parallel_forward(fn, batch2, fn_all=fn_all)
    ok = ok and torch.equal(got2, batch2 * 2) and calls == {"fn": 4, "all": 1} and D.plan_tile_waves(9, 8) == (8, 1)
    # stitched tiled result: disjoint regions gathered, not all-reduced
    boxes = [(0, 3, 0, 5), (0, 3, 5, 8), (3, 8, 0, 5), (3, 8, 5, 8), (8, 9, 0, 8)]   # (y0, y1, x0, x1): ragged sizes
    want = torch.arange(2 * 3 * 9 * 8, dtype=torch.float32).reshape(2, 3, 9, 8)
    mine = {i: want[:, :, b[0]:b[1], b[2]:b[3]].clone() for i, b in enumerate(boxes) if i % world == rank}
    res = D.gather_tile_regions(torch.zeros_like(want), mine, boxes)
    ok = ok and torch.equal(res, want)
    if rank == 0:
        out.put(bool(ok))
    torch.distributed.destroy_process_group()


def test_tile_parallel_groupnorm_reduction_and_batch_split_gloo_world2():
    """the only cross-GPU arithmetic of the tile-parallel single-image path: pixel-weighted GroupNorm statistics
    (GroupNormParam.summary semantics, infer/vaehook.py:489-508) reduced over ranks == the single-process value."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_gn, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    assert q.get(timeout=120) is True
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0


def test_shadow_ranks_and_padded_share():
    from omgsr_b200.dist import my_tiles, padded_share
    assert my_tiles(4, 5, 8) == ([1], True) and my_tiles(4, 3, 8) == ([3], False)
    assert my_tiles(9, 1, 8) == ([1], False) and my_tiles(9, 0, 8) == ([0, 8], False)
    idx, real = padded_share(9, 3, 8)
    assert idx == [3, 3] and real == 1
    idx, real = padded_share(9, 0, 8)
    assert idx == [0, 8] and real == 2
