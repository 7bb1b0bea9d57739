"""The GEGLU feed-forward GEMMs of one UNet evaluation (batch 1) timed back to back (host ahead, no events between
launches), plus the same shapes without the GEGLU epilogue for comparison. `ncu -k regex:gemm_tc -s 6 -c 1` on this
script captures the level-0 one (8192 x 2560 x 320)."""
import sys, ctypes as C, torch
sys.path.insert(0, ".")
import sdb200
from sdb200 import ops
from sdb200.unet import _pack_geglu
dev = torch.device("cuda:0")
lib = sdb200.lib.load()
def rec(M, C_, geglu):
    x = torch.randn(M, C_, device=dev).half()
    w = torch.randn(8 * C_, C_, device=dev) * 0.02; b = torch.randn(8 * C_, device=dev)
    if geglu:
        wp, bp = _pack_geglu(w, b)
        wp, bp = wp.to(dev), bp.to(dev)
    else:
        wp, bp = w.half().contiguous(), b
    ops.RECORD = []
    if geglu:
        ops.gemm(x, wp, bias=bp, act=ops.ACT_GEGLU, want_f16=True)
    else:
        ops.gemm(x, wp, bias=bp, want_f16=True, block_n=int(sys.argv[1]) if len(sys.argv) > 1 else 0)
    r, ops.RECORD = ops.RECORD, None
    return r[0]
def t(r, reps=100):
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for _ in range(3): lib.sdb_gemm(C.byref(r[0]), st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda._sleep(int(2e7)); e0.record()
    for _ in range(reps): lib.sdb_gemm(C.byref(r[0]), st)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1000
for M, C_ in ((8192, 320), (2048, 640), (512, 1280), (128, 1280)):
    a, b = rec(M, C_, True), rec(M, C_, False)
    fl = 2.0 * M * 8 * C_ * C_
    ta, tb = t(a), t(b)
    print(f"M={M:5d} N={8*C_:5d} K={C_:4d}: GEGLU {ta:6.1f} us ({fl/ta/1e6:5.0f} TF/s, bn={a[0].block_n})   plain fp16-out {tb:6.1f} us ({fl/tb/1e6:5.0f} TF/s, bn={b[0].block_n})")
