"""Image-parallel multi-GPU plumbing: one process per GPU, weights replicated, images dealt
round-robin, NO collective on the data path. torch.distributed (NCCL on GPUs, gloo in the CPU tests) is used
only for the timing barrier, the max-over-ranks reduction and an optional final gather of per-image results."""
import os

import torch
import torch.distributed as dist


def env_world():
    """(world_size, rank, local_rank) from the torchrun environment (defaults: single process)."""
    return (int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")),
            int(os.environ.get("LOCAL_RANK", "0")))


def init(backend=None, device=None):
    ws, rank, lr = env_world()
    if ws > 1 and not dist.is_initialized():
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        kw = {"device_id": device} if (backend == "nccl" and device is not None) else {}
        dist.init_process_group(backend, **kw)
    return ws, rank, lr


def shard_indices(n_items, rank, world):
    """round-robin deal of image indices: rank r gets r, r+world, ... (the reference's per-image loop,
    infer/infer_omgsr_s.py:69, split across replicas)."""
    return list(range(rank, n_items, world))


def max_over_ranks(value, device="cpu"):
    """max of a python float over all ranks (timing is reported as the slowest rank)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if t.is_cuda:
        t = t.float()
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value, device="cpu"):
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if t.is_cuda:
        t = t.float()
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def gather_results(local_items, n_items):
    """final gather of {index: tensor} dicts onto rank 0, restoring the original image order."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [local_items[i] for i in range(n_items)]
    bucket = [None] * dist.get_world_size()
    dist.all_gather_object(bucket, {k: v.cpu() for k, v in local_items.items()})
    if dist.get_rank() != 0:
        return None
    merged = {}
    for d in bucket:
        merged.update(d)
    return [merged[i] for i in range(n_items)]


def barrier():
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()
