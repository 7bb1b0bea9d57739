"""A/B of UNetModel.OVERLAP (skip 1x1 convs and V^T projections forked onto a side stream inside the captured graph):
graph-replay time of one guided SD-v1 evaluation with and without, same tuned tile choices, outputs compared.

usage: python scripts/overlap_ab.py [B]"""
import statistics
import sys

import torch

sys.path.insert(0, ".")
import sdb200
from sdb200 import arch

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
net = sdb200.UNetModel(**arch.SD_V1_UNET).load_weights(
    arch.random_state_dict(arch.unet_param_shapes(arch.SD_V1_UNET), 11, device=dev), dev)
g = torch.Generator(device=dev).manual_seed(0)
x = torch.randn(2 * B, 4, 64, 64, generator=g, device=dev)
ctx = torch.randn(2 * B, 77, 768, generator=g, device=dev)
t = torch.full((2 * B,), 981.0, device=dev)
net.use_cuda_graph = True
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def measure(overlap, n=30):
    net.OVERLAP = overlap
    net._graphs.clear()
    net._cap_stream = None
    eps = net(x, t, context=ctx).clone()
    gk = next(iter(net._graphs.values()))
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        gk["graph"].replay()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return statistics.median(ts), min(ts), gk["launches"], eps


res = {}
for rnd in range(2):
    for ov in (0, 1, 2):
        med, best, n, eps = measure(ov)
        res[ov] = eps
        print(f"round {rnd} overlap={ov}: median {med:.3f} ms, best {best:.3f} ms, {n} kernels", flush=True)
print("eps equal across modes:", torch.equal(res[1], res[0]), torch.equal(res[2], res[0]))
