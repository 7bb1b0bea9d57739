"""AutoencoderKL.decode at the bench size (64x64 latent -> 512x512): time with the tile model's choices, then with the
measured (autotuned) tile choices, and the per-shape GEMM breakdown.

usage: python scripts/vae_ms.py [latent]"""
import collections
import statistics
import sys

import torch

sys.path.insert(0, ".")
import sdb200
from sdb200 import arch, ops

dev = torch.device("cuda:0")
L = int(sys.argv[1]) if len(sys.argv) > 1 else 64
vae = sdb200.AutoencoderKL(**arch.SD_V1_VAE)
vae.load_weights(arch.random_state_dict(vae.shapes, 12, device=dev), dev)
z = torch.randn(1, 4, L, L, device=dev)


def timeit(n=10):
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        vae.decode(z, scale=1 / 0.18215, nhwc=True)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return statistics.median(ts)


def graph_time(n=10):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        vae.decode(z, scale=1 / 0.18215, nhwc=True)
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out = vae.decode(z, scale=1 / 0.18215, nhwc=True)
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return statistics.median(ts)


def breakdown(tag):
    ops.PROFILE = []
    vae.decode(z, scale=1 / 0.18215, nhwc=True)
    torch.cuda.synchronize()
    rec, ops.PROFILE = ops.PROFILE, None
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
    for kind, flops, e0, e1, shape in rec:
        a = agg[(kind,) + tuple(shape)]
        a[0] += 1
        a[1] += e0.elapsed_time(e1) * 1e3
        a[2] += flops
    tot = sum(a[1] for a in agg.values())
    print(f"--- {tag}: GEMM / attention launches {sum(a[0] for a in agg.values())}, {tot:.0f} us summed")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:12]:
        print(f"  {str(k):46s} n={a[0]:2d} {a[1]:8.1f} us  {a[2] / a[1] / 1e6:7.1f} TFLOP/s")


vae.decode(z, scale=1 / 0.18215, nhwc=True)
print(f"decode {8 * L}^2, tile model: eager {timeit():.3f} ms, graph {graph_time():.3f} ms", flush=True)
breakdown("tile model")
ops.AUTOTUNE = True
vae.decode(z, scale=1 / 0.18215, nhwc=True)
ops.AUTOTUNE = False
print(f"decode {8 * L}^2, autotuned:  eager {timeit():.3f} ms, graph {graph_time():.3f} ms", flush=True)
breakdown("autotuned")
