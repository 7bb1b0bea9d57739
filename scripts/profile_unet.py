"""One eager SD-v1 UNet evaluation (N_s=2, 64x64 latent) between cudaProfilerStart/Stop, for ncu launch lists."""
import sys, torch
sys.path.insert(0, ".")
import sdb200
from sdb200 import arch, ops
dev = torch.device("cuda:0")
ns = int(sys.argv[1]) if len(sys.argv) > 1 else 2
net = sdb200.UNetModel(**arch.SD_V1_UNET).load_weights(arch.random_state_dict(arch.unet_param_shapes(arch.SD_V1_UNET), 11, device=dev), dev)
g = torch.Generator(device=dev).manual_seed(0)
x = torch.randn(ns, 4, 64, 64, generator=g, device=dev)
ctx = torch.randn(ns, 77, 768, generator=g, device=dev)
t = torch.full((ns,), 981.0, device=dev)
kvs = net.set_context(ctx)
for _ in range(2):
    net._forward_impl(x, t, kvs)
torch.cuda.synchronize()
n0 = ops.launch_count()
torch.cuda.profiler.start()
net._forward_impl(x, t, kvs)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("launches in profiled eval:", ops.launch_count() - n0)
