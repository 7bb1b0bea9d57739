import sys, os, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests"); sys.path.insert(0, "oracle")
import sdb200
from sdb200 import ops
from helpers import golden, weights, CFGS, rel_l2
dev = torch.device("cuda:0")
case = golden("unet.pt")[int(os.environ.get("CASE", "0"))]
sd = weights("unet", case["cfg"], case["seed"])
x, t, ctx = case["x"].to(dev), case["t"].to(dev), case["ctx"].to(dev)
def mk(): return sdb200.UNetModel(**CFGS["unet"][case["cfg"]]).load_weights(sd, dev)
m = mk(); print("eager untuned      ", f"{rel_l2(m(x, t, context=ctx), case['eps']):.3e}")
m = mk(); m.use_cuda_graph = True; m.autotune = False
print("graph, no autotune ", f"{rel_l2(m(x, t, context=ctx), case['eps']):.3e}", " replay2", f"{rel_l2(m(x, t, context=ctx), case['eps']):.3e}")
m = mk(); m.use_cuda_graph = True; m.autotune = True
print("graph, autotune    ", f"{rel_l2(m(x, t, context=ctx), case['eps']):.3e}", " replay2", f"{rel_l2(m(x, t, context=ctx), case['eps']):.3e}")
print("tuned entries", len(ops.TUNED))
m = mk(); print("eager with TUNED   ", f"{rel_l2(m(x, t, context=ctx), case['eps']):.3e}")
bad = []
for k, v in ops.TUNED.items():
    pass
print(sorted(set(ops.TUNED.values())))
