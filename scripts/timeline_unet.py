"""Timeline of ONE graph-replayed SD-v1 UNet evaluation (N_s = 2B, 64x64 latent), two views:

1. CUPTI (torch.profiler) kernel activity records of a graph replay: per-kernel start / duration and the idle gap to
   the previous kernel -> gpurun_out/timeline_kernels.csv + a per-family summary (busy time vs gaps).
2. In-kernel phase stamps of every sdb_gemm launch (sdb_debug_trace): per launch the median over CTAs of
   setup / wait-for-predecessor / first operands landed / last MMA issued / accumulator ready / epilogue done, in
   SM clock cycles, plus the launch's wall span from %globaltimer -> gpurun_out/gemm_trace.txt

usage: python scripts/timeline_unet.py [B] [tag]
"""
import ctypes as C
import os
import statistics
import sys

import torch

sys.path.insert(0, ".")
import sdb200
from sdb200 import arch, ops

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
tag = sys.argv[2] if len(sys.argv) > 2 else "r02"
os.makedirs("gpurun_out", exist_ok=True)
net = sdb200.UNetModel(**arch.SD_V1_UNET).load_weights(
    arch.random_state_dict(arch.unet_param_shapes(arch.SD_V1_UNET), 11, device=dev), dev)
g = torch.Generator(device=dev).manual_seed(0)
x = torch.randn(2 * B, 4, 64, 64, generator=g, device=dev)
ctx = torch.randn(2 * B, 77, 768, generator=g, device=dev)
t = torch.full((2 * B,), 981.0, device=dev)
net.use_cuda_graph = True
net(x, t, context=ctx)          # autotune + capture
gk = next(iter(net._graphs.values()))
torch.cuda.synchronize()


def replay_ms(n=20):
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        gk["graph"].replay()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return statistics.median(ts)


base_ms = replay_ms()
print(f"graph replay: {base_ms:.3f} ms, {gk['launches']} kernels")

# ---------------------------------------------------------------- 1. CUPTI timeline
try:
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(3):
            gk["graph"].replay()
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and "memcpy" not in e.name.lower()
           and "memset" not in e.name.lower()]
    evs.sort(key=lambda e: e.time_range.start)
    n = len(evs) // 3
    evs = evs[2 * n:]          # last replay
    t0 = evs[0].time_range.start
    rows = []
    prev_end = t0
    for e in evs:
        s, d = e.time_range.start - t0, e.time_range.end - e.time_range.start
        rows.append((e.name.split("(")[0][:60], s, d, e.time_range.start - prev_end))
        prev_end = max(prev_end, e.time_range.end)
    span = prev_end - t0
    with open(f"gpurun_out/timeline_kernels_{tag}.csv", "w") as f:
        f.write("name,start_us,dur_us,gap_before_us\n")
        for r in rows:
            f.write(f"{r[0]},{r[1]:.2f},{r[2]:.2f},{r[3]:.2f}\n")
    fam = {}
    for name, s, d, gap in rows:
        k = name.replace("sdb::", "").split("<")[0]
        a = fam.setdefault(k, [0, 0.0, 0.0])
        a[0] += 1
        a[1] += d
        a[2] += max(gap, 0.0)
    with open(f"gpurun_out/timeline_summary_{tag}.txt", "w") as f:
        hdr = (f"one UNet evaluation under CUPTI tracing: span {span:.0f} us ({len(rows)} kernels); untraced graph replay "
               f"{base_ms * 1e3:.0f} us\nfamily                          n   busy_us  avg_us  gap_before_us(sum) avg_gap\n")
        f.write(hdr)
        print(hdr, end="")
        for k, a in sorted(fam.items(), key=lambda kv: -kv[1][1]):
            line = f"{k:30s} {a[0]:4d} {a[1]:9.1f} {a[1] / a[0]:7.2f} {a[2]:12.1f} {a[2] / a[0]:9.2f}\n"
            f.write(line)
            print(line, end="")
        tot_busy = sum(a[1] for a in fam.values())
        tot_gap = sum(a[2] for a in fam.values())
        line = f"total busy {tot_busy:.0f} us, total gaps {tot_gap:.0f} us (overlapping kernels count busy twice)\n"
        f.write(line)
        print(line, end="")
except Exception as ex:  # profiler unavailable: the in-kernel trace below still runs
    print("CUPTI timeline failed:", repr(ex))

# ---------------------------------------------------------------- 2. in-kernel phase stamps of the GEMM launches
lib = sdb200.lib.load()
SLOT = 8 + 8 * 160
buf = torch.zeros(SLOT * 720, dtype=torch.int64, device=dev)   # 3 passes (2 warm-up + capture) x ~210 launches
lib.sdb_debug_trace(C.c_void_p(buf.data_ptr()), buf.numel())
net._graphs = {}
net.autotune = False            # TUNED already holds the choices of the first capture
net(x, t, context=ctx)          # warm-up passes + capture with trace pointers baked into the nodes
gk2 = next(iter(net._graphs.values()))
used = lib.sdb_debug_trace(None, 0)
torch.cuda.synchronize()
n_launch_total = used // SLOT
per_pass = n_launch_total // 3  # two warm-up passes + the captured one
lib_ms = []
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    gk2["graph"].replay()
    e1.record()
    torch.cuda.synchronize()
    lib_ms.append(e0.elapsed_time(e1))
h = buf.cpu().view(-1, SLOT)[2 * per_pass: 3 * per_pass]
with open(f"gpurun_out/gemm_trace_{tag}.txt", "w") as f:
    f.write(f"traced graph replay {statistics.median(lib_ms):.3f} ms; {per_pass} gemm launches; cycles are SM clocks "
            "(median over CTAs), span/gap in us from %globaltimer\n")
    f.write("  #      M     N  kit tap  bn(p=pair) sp(c=cluster) grid | setup  wait   data  mmaend accrdy epiend |  span_us gap_us\n")
    prev_end = None
    tot_span = tot_gap = 0.0
    for i in range(per_pass):
        r = h[i]
        grid, bn, sp, kit, M, N, taps, tiles = [int(v) for v in r[:8]]
        cg, bn = bn // 1000, bn % 1000
        csk, sp = sp // 100, sp % 100
        if grid == 0:
            continue
        c = r[8: 8 + 8 * min(grid, 160)].view(-1, 8)
        gt0, gt1 = c[:, 0], c[:, 1]
        start, end = int(gt0.min()), int(gt1.max())
        med = [int(c[:, k].median()) for k in range(2, 8)]
        span = (end - start) / 1e3
        gap = (start - prev_end) / 1e3 if prev_end is not None else 0.0
        prev_end = end
        tot_span += span
        tot_gap += gap
        f.write(f"{i:3d} {M:6d} {N:5d} {kit:4d} {taps:3d} {bn:3d}{'p' if cg == 2 else ' '}{sp:2d}{'c' if csk else ' '}{grid:4d} | {med[0]:5d} {med[1]:5d} {med[2]:6d} "
                f"{med[3]:6d} {med[4]:6d} {med[5]:6d} | {span:8.2f} {gap:7.2f}\n")
    f.write(f"sum of gemm spans {tot_span:.0f} us, sum of gaps between gemm launches (other kernels + idle) {tot_gap:.0f} us\n")
print(open(f"gpurun_out/gemm_trace_{tag}.txt").read()[-3000:])
