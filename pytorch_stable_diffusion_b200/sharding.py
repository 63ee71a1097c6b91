"""Multi-GPU sampling: independent seeds sharded across ranks (one process per GPU).

The path has no cross-sample operation (GroupNorm, LayerNorm and attention are per sample; a CFG
pair stays on one GPU), so rank r of R owns a contiguous block of seeds, a full replica of the
weights and its own captured CUDA graph; there is no collective inside the loop. The only exchange
is an optional final gather of the uint8 images (786 KB each at 512x512).
"""
import torch
import torch.distributed as dist


def shard_range(n_items, rank, world_size):
    """[begin, end) of the items rank `rank` owns: contiguous blocks, sizes differing by at most one,
    earlier ranks taking the larger blocks."""
    if world_size <= 0 or not 0 <= rank < world_size:
        raise ValueError(f"bad rank {rank} / world size {world_size}")
    base, extra = divmod(n_items, world_size)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def shard_seeds(seeds, rank=None, world_size=None):
    """The seeds this rank samples."""
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    if world_size is None:
        world_size = dist.get_world_size() if dist.is_initialized() else 1
    b, e = shard_range(len(seeds), rank, world_size)
    return list(seeds[b:e])


def gather_images(local_images, n_total, group=None):
    """All-gathers per-rank uint8 image batches (B_r, H, W, 3) into (n_total, H, W, 3) in seed order.
    Works on the process group's backend (NCCL on GPUs, gloo on CPU)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local_images
    world = dist.get_world_size(group)
    t = local_images if torch.is_tensor(local_images) else torch.from_numpy(local_images)
    per = (n_total + world - 1) // world
    pad = torch.zeros((per,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    pad[:t.shape[0]] = t
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad, group=group)
    parts = []
    for r, o in enumerate(out):
        b, e = shard_range(n_total, r, world)
        parts.append(o[:e - b])
    return torch.cat(parts, 0)


def generate_sharded(prompt, uncond_prompt, seeds, gather=True, **kw):
    """pipeline.generate over `seeds`, sharded across the ranks of the default process group. Every
    rank returns all images when gather=True, else only its own shard. A rank whose shard is empty (more ranks
    than seeds) generates nothing and still takes part in the gather."""
    import numpy as np
    from . import pipeline
    mine = shard_seeds(seeds)
    h = kw.get("height") or pipeline.HEIGHT
    w = kw.get("width") or pipeline.WIDTH
    if mine:
        images = pipeline.generate(prompt, uncond_prompt, seeds=mine, batch_size=len(mine), return_all=True, **kw)
    else:
        images = np.zeros((0, h, w, 3), dtype=np.uint8)
    if not gather:
        return images
    t = torch.from_numpy(images)
    if dist.is_initialized() and dist.get_backend() == "nccl":
        dev = kw.get("device")
        t = t.to(torch.device("cuda", torch.cuda.current_device()) if dev is None else dev)
    out = gather_images(t, len(seeds))
    return out.cpu().numpy() if torch.is_tensor(out) else out
