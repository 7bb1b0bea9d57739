"""Graph-replay time of one guided SD-v1 UNet evaluation (N_s = 2B, 64x64 latent, L2 flushed between replays) and the
eps error against the reference golden when B = 1: the quick A/B number for kernel changes.

usage: python scripts/eval_ms.py [B]"""
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
net(x, t, context=ctx)
gk = next(iter(net._graphs.values()))
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
torch.cuda.synchronize()
ts = []
for _ in range(40):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    gk["graph"].replay()
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
print(f"UNet evaluation (graph replay, B={B}): median {statistics.median(ts):.3f} ms, best {min(ts):.3f} ms, "
      f"{gk['launches']} kernels", flush=True)
