import sys, torch, torch.nn.functional as F
sys.path.insert(0, ".")
import sdb200
from sdb200 import ops
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
def maxdiff(a, b): return float((a.double() - b.double()).abs().max())
# conv with stats, with/without splits
for (nb, h, w, c, n, sp) in [(2, 16, 16, 128, 128, 0), (2, 16, 16, 128, 128, 4), (2, 8, 8, 128, 256, -1), (2, 64, 64, 320, 320, 0)]:
    x = torch.randn(nb, h, w, c, generator=g).to(dev).half()
    wk = (torch.randn(n, 9 * c, generator=g) * 0.02).to(dev).half()
    res = torch.randn(nb * h * w, n, generator=g).to(dev)
    outs, sts = [], []
    for _ in range(4):
        _, o = ops.gemm(x, wk, taps=9, residual=res, want_f32=True, splits=sp, want_stats=True)
        outs.append(o.clone()); sts.append(ops.channel_stats(o).sum(0).clone())
    print("gemm", (nb, h, w, c, n, sp), "out diff", [maxdiff(o, outs[0]) for o in outs], "stats rel diff",
          [float(((s_ - sts[0]).abs() / sts[0].abs().clamp_min(1e-9)).max()) for s_ in sts])
    gam = torch.ones(n, device=dev); bet = torch.zeros(n, device=dev)
    ys = []
    for _ in range(3):
        _, o = ops.gemm(x, wk, taps=9, residual=res, want_f32=True, splits=sp, want_stats=True)
        y, _ = ops.groupnorm(o.view(nb, h, w, n), gam, bet, silu=True)
        ys.append(y.clone())
    print("   gn diff", [maxdiff(y, ys[0]) for y in ys])
# attention
q = torch.randn(2, 1024, 512, generator=g).to(dev).half(); k = torch.randn(2, 1024, 512, generator=g).to(dev).half()
vt = torch.randn(2, 512, 1024, generator=g).to(dev).half()
os_ = [ops.attention(q, k, vt, heads=8, d=40, dpad=64, nq=1024, nkv=1024, scale=40 ** -0.5).clone() for _ in range(4)]
print("attention diff", [maxdiff(o, os_[0]) for o in os_])
x = torch.randn(2048, 320, generator=g).to(dev)
ls = [ops.layernorm(x, torch.ones(320, device=dev), torch.zeros(320, device=dev)).clone() for _ in range(3)]
print("ln diff", [maxdiff(o, ls[0]) for o in ls])
# hi/lo gemm
a = torch.randn(2048, 320, generator=g).to(dev).half(); wq = (torch.randn(320, 960, generator=g) * 0.05).to(dev).half()
hs = [ops.gemm(a, wq, a1=a, a2=a, want_f32=True, splits=-1)[1].clone() for _ in range(3)]
print("hilo gemm diff", [maxdiff(o, hs[0]) for o in hs])
