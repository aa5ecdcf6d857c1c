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
