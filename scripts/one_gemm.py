import sys, torch
sys.path.insert(0, ".")
import sdb200
from sdb200 import ops
dev = torch.device("cuda:0")
M, N, K = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
bn = int(sys.argv[4]) if len(sys.argv) > 4 else 0
x = torch.randn(M, K, device=dev).half(); w = torch.randn(N, K, device=dev).half()
out = torch.empty(M, N, dtype=torch.float32, device=dev)
for _ in range(3):
    ops.gemm(x, w, out_f32=out, block_n=bn)
torch.cuda.synchronize()
