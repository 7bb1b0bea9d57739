"""Main-loop limiter probe: 3x3 conv 64x64x320 -> 320 (45 K-steps) and a large GEMM, per tile shape, with pieces of the
loop switched off through SDB_DBG (8: no B loads, 16: no MMAs, 32: no A loads). Prints cycles per K-step of the median
leader CTA (from sdb_debug_trace) and us / launch."""
import ctypes as C, os, statistics, sys, torch
sys.path.insert(0, ".")
import sdb200
from sdb200 import ops
dev = torch.device("cuda:0")
lib = sdb200.lib.load()
g = torch.Generator().manual_seed(0)
x4 = torch.randn(2, 64, 64, 320, generator=g).half().to(dev)
SLOT = 8 + 8 * 160
def run(name, fn, iters):
    ops.RECORD = []
    fn()
    rec, ops.RECORD = ops.RECORD[0], None
    d = rec[0]
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for _ in range(3): lib.sdb_gemm(C.byref(d), st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda._sleep(2_000_000); e0.record()
    for _ in range(10): lib.sdb_gemm(C.byref(d), st)
    e1.record(); torch.cuda.synchronize(); us = e0.elapsed_time(e1) / 10 * 1e3
    buf = torch.zeros(SLOT * 4, dtype=torch.int64, device=dev)
    lib.sdb_debug_trace(C.c_void_p(buf.data_ptr()), buf.numel())
    for _ in range(3): lib.sdb_gemm(C.byref(d), st)
    lib.sdb_debug_trace(None, 0); torch.cuda.synchronize()
    r = buf.cpu().view(-1, SLOT)[2]; grid = int(r[0]); c = r[8:8 + 8 * min(grid, 160)].view(-1, 8)
    lead = c[c[:, 4] > 0]          # CTAs that issued MMAs
    data, mmaend = int(lead[:, 4].median()), int(lead[:, 5].median())
    print(f"{name:34s} {us:8.2f} us | bn {int(r[1])} grid {grid} | first data {data:6d} last mma issued {mmaend:7d} -> "
          f"{(mmaend - data) / iters:7.1f} clk / K-step")
print("SDB_DBG =", os.environ.get("SDB_DBG", "0"))
for n in (320,):
    w9 = (torch.randn(n, 9 * 320, generator=g) * 0.02).half().to(dev)
    for bn, pair in ((160, 1), (160, 2), (128, 1), (128, 2), (64, 1), (32, 1)):
        run(f"conv3x3 8192x{n}x2880 bn{bn} pair{pair}", lambda: ops.gemm(x4, w9, taps=9, want_f32=True, block_n=bn, pair=pair), 45)
w9 = (torch.randn(1280, 9 * 320, generator=g) * 0.02).half().to(dev)
for bn, pair in ((256, 1), (256, 2), (160, 1), (160, 2), (128, 2)):
    run(f"conv3x3 8192x1280x2880 bn{bn} pair{pair}", lambda: ops.gemm(x4, w9, taps=9, want_f16=True, block_n=bn, pair=pair), 45 * (5 if bn == 256 else 8 if bn == 160 else 10) * (64 // pair) / (148 // pair) if False else 45)
