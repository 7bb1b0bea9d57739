import sys, torch, torch.nn.functional as F
sys.path.insert(0, ".")
import sdb200 as S
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(1)
nb, h, w, c, n = 2, 8, 8, 320, 1280
x = torch.randn(nb, h, w, c, generator=g).to(dev).half()
wk = (torch.randn(n, 9 * c, generator=g) * (9 * c) ** -0.5).to(dev).half()
wt = wk.reshape(n, 3, 3, c).permute(0, 3, 1, 2).float()
ref = F.conv2d(x.float().permute(0, 3, 1, 2), wt, padding=1).permute(0, 2, 3, 1).reshape(-1, n)
for bn in (0, 32, 64, 128, 256):
    for splits in (1, 2, 3, 5, 9):
        _, o = S.ops.gemm(x, wk, taps=9, want_f32=True, splits=splits, block_n=bn)
        torch.cuda.synchronize()
        e = (o - ref)
        print(f"bn={bn} splits={splits}: rel {float(e.norm()/ref.norm()):.3e}  max abs {float(e.abs().max()):.3e}",
              " bad cols:", int((e.abs().max(0).values > 1e-3).sum()), " bad rows:", int((e.abs().max(1).values > 1e-3).sum()))
# plain split
a = torch.randn(256, 1280, generator=g).to(dev).half()
b = (torch.randn(640, 1280, generator=g) * 1280 ** -0.5).to(dev).half()
r = a.float() @ b.float().t()
for splits in (1, 2, 4):
    _, o = S.ops.gemm(a, b, want_f32=True, splits=splits, block_n=128)
    print("plain splits", splits, float((o - r).norm() / r.norm()))
