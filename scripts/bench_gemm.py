"""Microbenchmark of sdb_gemm on the UNet's dominant shapes (CUDA events, back-to-back launches, L2 warm)."""
import sys, torch
sys.path.insert(0, ".")
import sdb200
from sdb200 import ops
dev = torch.device("cuda:0")
def t(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1000.0
def run(name, nb, h, w, c, n, taps, bns=(0,), splits=(0,)):
    x = torch.randn(nb, h, w, c, device=dev).half()
    wt = torch.randn(n, taps * c, device=dev).half()
    M = nb * h * w
    fl = 2.0 * M * n * taps * c
    out = torch.empty(M, n, dtype=torch.float32, device=dev)
    for bn in bns:
        for sp in splits:
            us = t(lambda: ops.gemm(x if taps == 9 else x.view(M, c), wt, taps=taps, out_f32=out, block_n=bn, splits=sp))
            print(f"{name:34s} M={M:5d} N={n:5d} K={taps*c:6d} bn={bn:3d} splits={sp:2d}: {us:7.1f} us  {fl/us/1e6:7.1f} TFLOP/s")
run("conv3 L0 320->320", 2, 64, 64, 320, 320, 9, bns=(0, 64, 128, 160, 256))
run("plain same K (2880)", 2, 64, 64, 2880, 320, 1, bns=(160,))
run("conv3 L0 640cat->320", 2, 64, 64, 640, 320, 9, bns=(160,))
run("conv3 L1 640->640", 2, 32, 32, 640, 640, 9, bns=(0, 128, 160), splits=(0, -1, 2, 3))
run("conv3 L2 1280->1280", 2, 16, 16, 1280, 1280, 9, bns=(0, 128, 256), splits=(-1, 4, 8, 16))
run("conv3 L3 1280->1280 (8x8)", 2, 8, 8, 1280, 1280, 9, bns=(0, 64, 128), splits=(-1, 8, 16, 32))
run("linear L0 320->320", 2, 64, 64, 320, 320, 1, bns=(0, 64, 160))
run("linear qk L0 320->1024", 2, 64, 64, 320, 1024, 1, bns=(0, 128, 256))
run("geglu-size L0 320->2560", 2, 64, 64, 320, 2560, 1, bns=(0, 128, 256))
run("ff2 L0 1280->320", 2, 64, 64, 1280, 320, 1, bns=(0, 160))
run("1x1 hilo L0 960->320", 2, 64, 64, 960, 320, 1, bns=(0, 160))
run("big square 8192^2 x 4096", 1, 1, 8192, 4096, 8192, 1, bns=(128, 256))
