import sys, torch, torch.nn.functional as F
sys.path.insert(0, ".")
import sdb200 as S
dev = torch.device("cuda:0")
def run(nb, h, w, c, n, bn=0):
    g = torch.Generator().manual_seed(1)
    x = torch.randn(nb, h, w, c, generator=g).to(dev).half()
    wt = (torch.randn(n, c, 3, 3, generator=g) * (9 * c) ** -0.5).to(dev).half()
    wk = wt.permute(0, 2, 3, 1).reshape(n, 9 * c).contiguous()
    _, o = S.ops.gemm(x, wk, taps=9, want_f32=True, block_n=bn)
    torch.cuda.synchronize()
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), wt.float(), padding=1).permute(0, 2, 3, 1)
    o = o.reshape(nb, h, w, n)
    err = (o - ref).pow(2).sum(-1).sqrt() / ref.pow(2).sum(-1).sqrt()
    print(f"conv nb={nb} h={h} w={w} c={c} n={n}: total rel {float((o-ref).norm()/ref.norm()):.3e}")
    bad = err > 1e-3
    print("  bad rows per sample/y:", [[int(bad[i, y].sum()) for y in range(h)] for i in range(nb)])
    # single-tap probes: which taps are wrong?
    for tap in range(9):
        wt1 = torch.zeros_like(wt); wt1[:, :, tap // 3, tap % 3] = wt[:, :, tap // 3, tap % 3]
        wk1 = wt1.permute(0, 2, 3, 1).reshape(n, 9 * c).contiguous()
        _, o1 = S.ops.gemm(x, wk1, taps=9, want_f32=True, block_n=bn)
        r1 = F.conv2d(x.float().permute(0, 3, 1, 2), wt1.float(), padding=1).permute(0, 2, 3, 1)
        e = (o1.reshape(nb, h, w, n) - r1).pow(2).sum(-1).sqrt()
        print(f"   tap {tap}: bad px per sample {[int((e[i] > 1e-2).sum()) for i in range(nb)]}", end="")
        if int((e > 1e-2).sum()):
            ys, xs = torch.nonzero(e[0] > 1e-2, as_tuple=True)
            print("  sample0 bad y range", int(ys.min()) if len(ys) else None, int(ys.max()) if len(ys) else None,
                  "x range", int(xs.min()) if len(xs) else None, int(xs.max()) if len(xs) else None, end="")
        print()
run(2, 16, 16, 64, 128)
run(1, 16, 16, 64, 64)
run(2, 8, 8, 64, 64)
