import sys, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import sdb200
from sdb200 import ops
from helpers import golden, weights, CFGS
dev = torch.device("cuda:0")
case = golden("unet.pt")[2]
m = sdb200.UNetModel(**CFGS["unet"]["sdv1"]).load_weights(weights("unet", "sdv1", case["seed"]), dev)
x, t, ctx = case["x"].to(dev), case["t"].to(dev), case["ctx"].to(dev)
m.set_context(ctx)
trace = []
names = ["gemm", "groupnorm", "layernorm", "attention"]
orig = {n: getattr(ops, n) for n in names}
def wrap(n):
    def f(*a, **k):
        r = orig[n](*a, **k)
        outs = r if isinstance(r, tuple) else (r,)
        info = ""
        if n == "gemm":
            info = f"M={a[0].numel()//a[0].shape[-1]} n={a[1].shape[0]} K={a[1].shape[1]} taps={k.get('taps',1)} stats={k.get('want_stats',False)}"
        trace[-1].append((n, info, [o.detach().float().clone() for o in outs if torch.is_tensor(o)]))
        return r
    return f
for n in names: setattr(ops, n, wrap(n))
for run in range(2):
    trace.append([])
    m(x, t, context=ctx)
torch.cuda.synchronize()
A, B = trace
print(len(A), len(B))
shown = 0
for i, (ra, rb) in enumerate(zip(A, B)):
    d = max(float((u - v).abs().max()) for u, v in zip(ra[2], rb[2]))
    mag = max(float(u.abs().max()) for u in ra[2])
    nd = sum(int(((u - v).abs() > 0).sum()) for u, v in zip(ra[2], rb[2]))
    tot = sum(u.numel() for u in ra[2])
    if d > 0 and shown < 25:
        print(f"op {i:3d} {ra[0]:10s} {ra[1]:60s} maxdiff {d:.3e} (max {mag:.2f}) differing {nd}/{tot}")
        shown += 1
