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


# ---- tile-parallel single-image path (reference tile semantics) ---------------------------------
TILE_PARALLEL = False  # set by enable_tile_parallel(); read by infer/vaehook.py and the _forward_tile methods


def enable_tile_parallel(on=True):
    """spread the latent tiles (denoiser) and the VAE tiles of ONE image over all ranks of the default group."""
    global TILE_PARALLEL
    TILE_PARALLEL = bool(on)


def tile_parallel():
    return TILE_PARALLEL and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def broadcast_(t, src=0):
    if tile_parallel():
        dist.broadcast(t, src=src)
    return t


def plan_tile_waves(n_items, world):
    """(items handled data-parallel in full waves, left-over items). 9 latent tiles on 8 ranks -> (8, 1): the first eight
    go one per rank, the ninth is run by ALL ranks together (sequence-parallel) instead of costing a second wave."""
    full = (n_items // world) * world
    return full, n_items - full


def tile_parallel_forward(fn, batch, fn_all=None):
    """run `fn` on this rank's share of `batch` (dim 0 = independent tiles) and all-gather the results back into
    the original order. fn must map [n, ...] -> [n, ...].
    fn_all (optional): a COOPERATIVE form of fn that every rank calls with the same [1, ...] item and that returns the
    full result on every rank (the FLUX engine's sequence-parallel forward). With it, items that do not fill a wave
    are not padded into one more data-parallel wave (9 tiles on 8 GPUs = 2 waves = 4.5x at best) but computed by all
    ranks together."""
    world, rank = dist.get_world_size(), dist.get_rank()
    n = batch.shape[0]
    full, rest = plan_tile_waves(n, world) if fn_all is not None else (n, 0)
    outs = [None] * n
    if full:
        idx, _ = padded_share(full, rank, world)
        out = fn(batch[idx].contiguous()).contiguous()
        g = all_gather_batch(out)  # [world, per, ...]
        for i in range(full):
            outs[i] = g[i % world, i // world]
    for i in range(full, full + rest):
        outs[i] = fn_all(batch[i:i + 1].contiguous())[0]
    return torch.stack(outs, 0)


def reduce_gn_partials(num, wsum, group=None):
    """cross-GPU step of the tiled VAE's GroupNorm: every rank holds the pixel-weighted numerators of ITS tiles
    (sum_e pix_e * (mean_e, var_e)) and the weights (sum_e pix_e); the shared statistics are the all-reduced ratio.
    One tiny all-reduce ([n_img, 32, 2] + [n_img] floats) per GroupNorm site; nothing else crosses NVLink."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        packed = torch.cat([num.reshape(-1), wsum.reshape(-1)])
        dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
        num = packed[: num.numel()].reshape(num.shape)
        wsum = packed[num.numel():]
    return num / wsum.clamp_min(1e-30).reshape(-1, 1, 1)


def my_tiles(n_tiles, rank, world):
    """tile indices owned by `rank` (round robin) and whether the rank only shadows a tile (world > n_tiles):
    a shadow rank runs tile rank % n_tiles with weight 0 so that it stays in lock step with the collectives."""
    own = list(range(rank, n_tiles, world))
    if own:
        return own, False
    return [rank % n_tiles], True


def padded_share(n_items, rank, world):
    """equal-length index lists for an all_gather over the batch axis: rank r gets items r, r+world, ... padded
    with its last item (or item 0) up to ceil(n_items / world). Returns (indices, number of real entries)."""
    per = (n_items + world - 1) // world
    own = list(range(rank, n_items, world))
    real = len(own)
    while len(own) < per:
        own.append(own[-1] if own else 0)
    return own, real


def all_gather_batch(t, group=None):
    """equal-shaped per-rank tensors -> [world, ...] (rank-major), one collective into one buffer."""
    world = dist.get_world_size(group)
    t = t.contiguous()
    out = torch.empty((world,) + tuple(t.shape), device=t.device, dtype=t.dtype)
    if t.is_cuda:
        dist.all_gather_into_tensor(out, t, group=group)
    else:  # gloo (CPU tests) has no all_gather_into_tensor
        dist.all_gather(list(out.unbind(0)), t, group=group)
    return out


def gather_tile_regions(result, regions, n_tiles, group=None):
    """stitch a tiled result whose tiles were computed on different ranks WITHOUT reducing the whole image: every rank
    contributes the valid regions of its own tiles (padded to one shape), one all-gather moves them, every rank pastes
    all of them. `regions`: {tile index: (crop tensor [N, C, h, w], (y0, y1, x0, x1))} for the tiles this rank owns
    (round robin: tile i belongs to rank i % world); shapes[i] = (h, w) of every tile's region must be derivable by all
    ranks, so the caller passes the full list of boxes through `n_tiles` = [(y0, y1, x0, x1)] * n."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    boxes = n_tiles
    per = (len(boxes) + world - 1) // world
    mh = max(b[1] - b[0] for b in boxes)
    mw = max(b[3] - b[2] for b in boxes)
    N, C = result.shape[0], result.shape[1]
    send = torch.zeros((per, N, C, mh, mw), device=result.device, dtype=result.dtype)
    for k, i in enumerate(range(rank, len(boxes), world)):
        if i in regions:
            crop = regions[i]
            send[k, :, :, :crop.shape[2], :crop.shape[3]] = crop
    recv = all_gather_batch(send, group)  # [world, per, N, C, mh, mw]
    for r in range(world):
        for k, i in enumerate(range(r, len(boxes), world)):
            y0, y1, x0, x1 = boxes[i]
            result[:, :, y0:y1, x0:x1] = recv[r, k, :, :, :y1 - y0, :x1 - x0]
    return result
