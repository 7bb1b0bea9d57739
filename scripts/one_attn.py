import sys, torch
sys.path.insert(0, ".")
import sdb200
from sdb200 import ops
dev = torch.device("cuda:0")
B, H, N, d, dp = 2, 8, 4096, 40, 64
q = torch.randn(B, N, H * dp, device=dev).half(); k = torch.randn(B, N, H * dp, device=dev).half()
vt = torch.randn(B, H * dp, N, device=dev).half()
for _ in range(3):
    ops.attention(q, k, vt, heads=H, d=d, dpad=dp, nq=N, nkv=N, scale=d ** -0.5)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    ops.attention(q, k, vt, heads=H, d=d, dpad=dp, nq=N, nkv=N, scale=d ** -0.5)
e1.record(); torch.cuda.synchronize()
print("attention N=4096 d=40 B=2:", e0.elapsed_time(e1) / 10 * 1000, "us")
