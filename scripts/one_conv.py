import sys, torch
sys.path.insert(0, ".")
import sdb200
from sdb200 import ops
dev = torch.device("cuda:0")
x = torch.randn(2, 64, 64, 320, device=dev).half(); w = (torch.randn(320, 2880, device=dev) * 0.02).half()
bias = torch.randn(320, device=dev); res = torch.randn(8192, 320, device=dev)
for _ in range(3):
    ops.gemm(x, w, taps=9, bias=bias, residual=res, want_f32=True, want_stats=True)
torch.cuda.synchronize()
