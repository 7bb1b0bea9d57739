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


class CFGParallel:
    """Latency mode: the two halves of a classifier-free-guided evaluation (plms.py:182-186,
    ddim.py:174-178) run on TWO GPUs — group rank 0 evaluates the unconditional batch, rank 1 the conditional one —
    and the eps halves meet in the fused sampler-step kernel, which both ranks execute identically (bit-exact
    arithmetic), so the latents stay in lock-step without any further traffic.

    mode "p2p":  each rank publishes its eps into a symmetric (peer-mapped) buffer, a device-side barrier follows, and
                 the step kernel itself loads the other half straight from the peer GPU over NVLink (`eps_cond` /
                 `eps2` of sdb_sampler_step point into peer memory): the exchange is fused into the compute kernel.
                 Two alternating slots make the publish of evaluation k+1 safe against a late reader of evaluation k.
    mode "nccl": all_gather_into_tensor of the two halves, then the ordinary step kernel (works with gloo on CPU
                 tensors too: the host-logic test).
    """

    def __init__(self, group=None, mode="p2p", device=None, max_numel=8 * 4 * 96 * 96, barrier_timeout_ms=20000):
        self.group = group if group is not None else dist.group.WORLD
        assert dist.get_world_size(self.group) == 2, "CFG-parallel pairs exactly two ranks"
        self.role = dist.get_rank(self.group)          # 0: unconditional half, 1: conditional half
        self.mode = mode
        self.device = device
        self.evals = 0
        self._timeout = barrier_timeout_ms
        self._gather = {}
        if mode == "p2p":
            import torch.distributed._symmetric_memory as symm_mem
            self._pub = symm_mem.empty((2, max_numel), dtype=torch.float32, device=device)
            self._hdl = symm_mem.rendezvous(self._pub, self.group)
            self._peer = self._hdl.get_buffer(1 - self.role, (2, max_numel), torch.float32)
            self.max_numel = max_numel
        elif mode != "nccl":
            raise ValueError(f"unknown CFG-parallel mode {mode!r}")

    def select(self, uncond, cond):
        """This rank's half of the conditioning."""
        return uncond if self.role == 0 else cond

    def exchange(self, eps_local):
        """eps of this rank's half -> (eps_uncond_or_pair, eps_cond_or_None) for ops.sampler_step / dpm_solver_step."""
        n = eps_local.numel()
        k = self.evals
        self.evals += 1
        if self.mode == "nccl":
            buf = self._gather.get((n, eps_local.device))
            if buf is None:
                buf = self._gather[(n, eps_local.device)] = torch.empty((2,) + tuple(eps_local.shape), dtype=eps_local.dtype,
                                                                        device=eps_local.device)
            dist.all_gather_into_tensor(buf.view(-1), eps_local.contiguous().view(-1), group=self.group)
            return buf.flatten(0, 1), None
        assert n <= self.max_numel, (n, self.max_numel)
        slot = k & 1
        self._pub[slot, :n].copy_(eps_local.reshape(-1))            # publish (local write)
        self._hdl.barrier(channel=slot, timeout_ms=self._timeout)    # both halves published and visible
        mine = self._pub[slot, :n].view(eps_local.shape)
        peer = self._peer[slot, :n].view(eps_local.shape)            # peer GPU memory, read by the step kernel
        return (mine, peer) if self.role == 0 else (peer, mine)
