"""Data-parallel sharding of prompt batches over the GPUs of one box: each sample's trajectory is
independent (no cross-sample op in UNet / VAE / CLIP), so ranks are replicas with no per-step communication.
Collectives (torch.distributed: NCCL over NVLink on GPUs, gloo in the CPU tests):
  - broadcast_weights: one-time broadcast of the packed weight tensors from rank 0 instead of N host loads;
  - gather_images: decoded uint8 images to rank 0 once per batch.
Per-sample inputs (start noise) are derived from the GLOBAL sample index, so results do not depend on world size."""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int):
    """Contiguous shard [lo, hi) of n_items for `rank`; earlier ranks take the remainder."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def noise_for(global_index: int, shape, seed: int = 42):
    """Start noise x_T of one sample, a function of (seed, global sample index) only (CPU generator: identical on
    every rank and for every world size; the reference draws the whole batch from the device RNG, plms.py:124)."""
    g = torch.Generator().manual_seed(seed * 1_000_003 + int(global_index))
    return torch.randn(tuple(shape), generator=g)


def batch_noise(lo: int, hi: int, shape, seed: int = 42):
    return torch.stack([noise_for(i, shape, seed) for i in range(lo, hi)]) if hi > lo else torch.empty((0, *shape))


def _walk(o, out, seen):
    if torch.is_tensor(o):
        if o.data_ptr() not in seen:
            seen.add(o.data_ptr())
            out.append(o)
    elif isinstance(o, dict):
        for v in o.values():
            _walk(v, out, seen)
    elif isinstance(o, (list, tuple)):
        for v in o:
            _walk(v, out, seen)


def weight_tensors(*packed):
    """Unique tensors of the packed weight dictionaries (UNetModel.W, AutoencoderKL.W, FrozenCLIPEmbedder.W)."""
    out, seen = [], set()
    for w in packed:
        _walk(w, out, seen)
    return out


def broadcast_weights(*packed, src: int = 0):
    """Broadcast every packed weight tensor from `src` (in place). Returns the number of bytes sent."""
    n = 0
    for t in weight_tensors(*packed):
        dist.broadcast(t, src=src)
        n += t.numel() * t.element_size()
    return n


def gather_images(img: torch.Tensor, dst: int = 0):
    """Gather equally-shaped uint8 image batches to `dst`; returns [world*B, H, W, 3] there (rank order = global
    sample order for equal shards), None elsewhere."""
    world = dist.get_world_size()
    rank = dist.get_rank()
    bufs = [torch.empty_like(img) for _ in range(world)] if rank == dst else None
    dist.gather(img, bufs, dst=dst)
    return torch.cat(bufs, 0) if rank == dst else None
