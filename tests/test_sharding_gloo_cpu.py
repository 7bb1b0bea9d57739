"""N>1 path on CPU: two gloo ranks shard a prompt batch, broadcast 'weights', run a stand-in step and gather images.
The rank-gathered result must equal the single-process result bit for bit."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import sdb200


def _fake_pipeline(weights, noise, ids):
    """Deterministic stand-in for one denoise+decode (the real one needs a GPU): per-sample, no cross-sample op."""
    x = noise.flatten(1) @ weights["w"] + weights["b"] + ids.float().mean(1, keepdim=True)
    return (torch.tanh(x).reshape(-1, 4, 4, 3) * 127 + 128).clamp(0, 255).to(torch.uint8)


def _worker(rank, world, port, n_total, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    D = sdb200.dist
    g = torch.Generator().manual_seed(7)
    weights = {"w": torch.randn(4 * 8 * 8, 48, generator=g), "b": [torch.randn(48, generator=g)]}
    if rank != 0:                                  # non-root ranks start from garbage and must receive rank 0's
        weights = {"w": torch.zeros(4 * 8 * 8, 48), "b": [torch.zeros(48)]}
    nbytes = D.broadcast_weights(weights, src=0)
    assert nbytes == (4 * 8 * 8 * 48 + 48) * 4
    lo, hi = D.shard_range(n_total, rank, world)
    ids = torch.arange(n_total * 5).reshape(n_total, 5)[lo:hi]
    img = _fake_pipeline({"w": weights["w"], "b": weights["b"][0]}, D.batch_noise(lo, hi, (4, 8, 8), seed=42), ids)
    out = D.gather_images(img, dst=0)
    if rank == 0:
        ret.put(out)
    else:
        assert out is None
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_matches_single_process():
    n_total = 6
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_total, ret)) for r in range(2)]
    [p.start() for p in procs]
    gathered = ret.get(timeout=120)
    [p.join(60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    g = torch.Generator().manual_seed(7)
    w = {"w": torch.randn(4 * 8 * 8, 48, generator=g), "b": torch.randn(48, generator=g)}
    ids = torch.arange(n_total * 5).reshape(n_total, 5)
    single = _fake_pipeline(w, sdb200.dist.batch_noise(0, n_total, (4, 8, 8), seed=42), ids)
    assert gathered.shape == (n_total, 4, 4, 3) and torch.equal(gathered, single)


def _cfg_worker(rank, world, port, ret):
    """CFG-parallel host logic (dist.CFGParallel, mode 'nccl' = all_gather; gloo here): each rank evaluates its half,
    the exchange returns [e_uncond; e_cond] in that order on BOTH ranks, and a guided update computed from it is
    identical on the two ranks and equal to the single-process update."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cp = sdb200.dist.CFGParallel(mode="nccl")
    assert cp.role == rank
    g = torch.Generator().manual_seed(11)
    x = torch.randn(3, 4, 8, 8, generator=g)
    uc, c = torch.randn(3, 77, 16, generator=g), torch.randn(3, 77, 16, generator=g)
    model = lambda x_, cc: torch.tanh(x_) * cc.mean(dim=(1, 2))[:, None, None, None]   # noqa: E731
    mine = cp.select(uc, c)
    assert torch.equal(mine, uc if rank == 0 else c)
    outs = []
    for step in range(3):                      # repeated exchanges reuse the gather buffer
        eps2, eps_cond = cp.exchange(model(x + step, mine))
        assert eps_cond is None and eps2.shape == (6, 4, 8, 8)
        e_u, e_c = eps2[:3], eps2[3:]
        outs.append((e_u + 7.5 * (e_c - e_u)).clone())
    ret.put((rank, torch.stack(outs)))
    dist.barrier()
    dist.destroy_process_group()


def test_cfg_parallel_exchange_orders_halves():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    procs = [ctx.Process(target=_cfg_worker, args=(r, 2, port, ret)) for r in range(2)]
    [p.start() for p in procs]
    got = dict(ret.get(timeout=120) for _ in range(2))
    [p.join(60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    g = torch.Generator().manual_seed(11)
    x = torch.randn(3, 4, 8, 8, generator=g)
    uc, c = torch.randn(3, 77, 16, generator=g), torch.randn(3, 77, 16, generator=g)
    model = lambda x_, cc: torch.tanh(x_) * cc.mean(dim=(1, 2))[:, None, None, None]   # noqa: E731
    single = torch.stack([model(x + s, uc) + 7.5 * (model(x + s, c) - model(x + s, uc)) for s in range(3)])
    assert torch.equal(got[0], got[1]) and torch.equal(got[0], single)
