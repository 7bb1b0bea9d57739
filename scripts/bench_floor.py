"""Per-launch floor of the GEMM kernel: replay prebuilt descriptors back to back through the C ABI (host ahead of GPU)."""
import sys, ctypes as C, torch
sys.path.insert(0, ".")
import sdb200
from sdb200 import ops
dev = torch.device("cuda:0")
lib = sdb200.lib.load()
def rec(M, N, K, taps=1, hw=None, **kw):
    if taps == 9:
        nb, h, w, c = hw
        x = torch.randn(nb, h, w, c, device=dev).half(); wt = torch.randn(N, 9 * c, device=dev).half()
    else:
        x = torch.randn(M, K, device=dev).half(); wt = torch.randn(N, K, device=dev).half()
    ops.RECORD = []
    ops.gemm(x, wt, taps=taps, want_f32=True, **kw)
    r, ops.RECORD = ops.RECORD, None
    return r[0]
def t(recs, reps=200):
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for d, _, _ in recs: lib.sdb_gemm(C.byref(d), st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda._sleep(int(4e7))
    e0.record()
    for _ in range(reps):
        for d, _, _ in recs: lib.sdb_gemm(C.byref(d), st)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps / len(recs) * 1000
for name, r in [("tiny 128x64x64", rec(128, 64, 64)),
                ("1 tile 128x128x320", rec(128, 128, 320)),
                ("148 tiles 128x128 K=320", rec(128 * 148, 128, 320)),
                ("linear 8192x320x320", rec(8192, 320, 320)),
                ("linear 8192x320x320 +stats", rec(8192, 320, 320, rows_per_sample=4096, want_stats=True)),
                ("linear 512x1280x1280", rec(512, 1280, 1280, splits=-1)),
                ("linear 512x1280x1280 no split", rec(512, 1280, 1280)),
                ("conv 8x8 1280->1280 auto split", rec(0, 1280, 0, taps=9, hw=(2, 8, 8, 1280), splits=-1)),
                ("conv 16x16 1280->1280 auto split", rec(0, 1280, 0, taps=9, hw=(2, 16, 16, 1280), splits=-1)),
                ("conv 64x64 320->320", rec(0, 320, 0, taps=9, hw=(2, 64, 64, 320)))]:
    print(f"{name:36s} {t([r]):7.2f} us/launch   (bn={r[0].block_n}, splits={r[0].splits})")
x = torch.randn(8192, 320, device=dev); g = torch.ones(320, device=dev); b = torch.zeros(320, device=dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda._sleep(int(4e7)); e0.record()
for _ in range(200): ops.layernorm(x, g, b)
e1.record(); torch.cuda.synchronize(); print("layernorm 8192x320", e0.elapsed_time(e1) / 200 * 1000, "us (python-launch bound if > kernel)")
# fused-statistics cost without the memset (arena path)
ops.ARENA = ops.StatsArena(dev)
r = rec(8192, 320, 320, rows_per_sample=4096, want_stats=True)
print(f"{'linear 8192x320x320 +stats (arena)':36s} {t([r]):7.2f} us/launch")
ops.ARENA.reset()
r = rec(0, 320, 0, taps=9, hw=(2, 64, 64, 320), want_stats=True)
print(f"{'conv 64x64 320->320 +stats (arena)':36s} {t([r]):7.2f} us/launch")
ops.ARENA = None
