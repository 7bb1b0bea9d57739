"""Where the GEMM time of one UNet evaluation goes (batch 1, N_s=2): record every sdb_gemm descriptor with the tuned
tiles, replay the whole sequence in order through the C ABI with an event pair around every launch (host ahead of the
GPU), and print per-shape totals against the per-launch ideal max(flops/peak, bytes/HBM)."""
import sys, ctypes as C, collections, torch
sys.path.insert(0, ".")
import sdb200
from sdb200 import ops, arch
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
unet = sdb200.UNetModel(**arch.SD_V1_UNET).load_weights(arch.random_state_dict(arch.unet_param_shapes(arch.SD_V1_UNET), 11, device=dev), dev)
x = torch.randn(2 * B, 4, 64, 64, device=dev); t = torch.full((2 * B,), 981.0, device=dev)
ctx = torch.randn(2 * B, 77, 768, device=dev)
unet.use_cuda_graph = True
unet(x, t, context=ctx)                      # autotune + capture
unet.use_cuda_graph = False
kvs = unet.context_kv(ctx)
unet._forward_impl(x, t, kvs)
ops.RECORD = []
keep = unet._forward_impl(x, t, kvs)
recs, ops.RECORD = ops.RECORD, None
lib = sdb200.lib.load(); st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
n = len(recs); REPS = 15
ev = [[torch.cuda.Event(enable_timing=True) for _ in range(n + 1)] for _ in range(REPS)]
for r in range(REPS):
    torch.cuda._sleep(int(4e7))
    ev[r][0].record()
    for i, (d, _, _) in enumerate(recs):
        lib.sdb_gemm(C.byref(d), st)
        ev[r][i + 1].record()
torch.cuda.synchronize()
us = [sorted(ev[r][i].elapsed_time(ev[r][i + 1]) * 1e3 for r in range(REPS))[REPS // 2] for i in range(n)]
rows = collections.OrderedDict()
for (d, fl, keepalive), u in zip(recs, us):
    M = d.nb * d.h * d.w; K = d.taps * (d.c0 + d.c1 + d.c2 + d.c3)
    byt = 2.0 * (M * (d.c0 + d.c1 + d.c2 + d.c3) + d.n * K) + (4.0 if d.out_f32 else 2.0) * M * d.n
    ideal = max(2.0 * M * d.n * K / 1460e12, byt / 7.0e12) * 1e6
    key = (M, d.n, K, d.taps, d.block_n, d.splits, bool(d.stats_out), bool(d.out_f16_lo))
    e = rows.setdefault(key, [0, 0.0, 0.0, 0.0]); e[0] += 1; e[1] += u; e[2] += ideal; e[3] += 2.0 * M * d.n * K
tot = sum(us)
print(f"{n} gemm launches, {tot:.0f} us in sequence (event pairs add ~1 us each)")
print(f"{'M':>6} {'N':>6} {'K':>6} tap  bn  sp st lo |   n   total_us  avg_us  ideal_us  TF/s  share")
for k, e in sorted(rows.items(), key=lambda kv: -kv[1][1]):
    M, N, K, taps, bn, sp, stt, lo = k
    print(f"{M:6d} {N:6d} {K:6d} {taps:3d} {bn:3d} {sp:3d} {int(stt):2d} {int(lo):2d} | {e[0]:3d} {e[1]:9.1f} {e[1]/e[0]:7.1f} {e[2]/e[0]:8.1f} {e[3]/e[1]/1e6:6.0f} {100*e[1]/tot:5.1f}%")
lv = collections.OrderedDict()
for k, e in rows.items():
    a = lv.setdefault(k[0], [0, 0.0, 0.0]); a[0] += e[0]; a[1] += e[1]; a[2] += e[2]
print("by M (rows):")
for M, a in sorted(lv.items(), key=lambda kv: -kv[0]):
    print(f"  M={M:6d}  n={a[0]:3d}  {a[1]:8.1f} us  ideal {a[2]:7.1f} us")
