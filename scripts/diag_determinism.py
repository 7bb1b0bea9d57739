import sys, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import sdb200
from helpers import golden, weights, CFGS, rel_l2
dev = torch.device("cuda:0")
for idx in (0, 2):
    case = golden("unet.pt")[idx]
    m = sdb200.UNetModel(**CFGS["unet"][case["cfg"]]).load_weights(weights("unet", case["cfg"], case["seed"]), dev)
    x, t, ctx = case["x"].to(dev), case["t"].to(dev), case["ctx"].to(dev)
    outs = [m(x, t, context=ctx).clone() for _ in range(4)]
    m.set_context(ctx)
    outs += [m(x, t, context=ctx).clone() for _ in range(3)]
    torch.cuda.synchronize()
    print(case["cfg"], tuple(x.shape), "vs ref", [f"{rel_l2(o, case['eps']):.2e}" for o in outs])
    print("   vs run0", [f"{rel_l2(o, outs[0]):.2e}" for o in outs])
