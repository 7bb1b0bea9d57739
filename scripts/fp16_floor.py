"""Estimate the fp16-operand noise floor of the UNet on CPU by rounding GEMM operands in the oracle."""
import sys, torch, types
sys.path.insert(0, "."); sys.path.insert(0, "oracle"); sys.path.insert(0, "tests")
import torch.nn.functional as F
import ldm_oracle as O
from helpers import CFGS, golden, rel_l2, weights

def run(case, round_w, round_a, round_attn):
    sd = weights("unet", case["cfg"], case["seed"])
    q = lambda t: t.half().float()
    if round_w:
        sd = {k: (q(v) if (v.dim() >= 2 and "time_embed" not in k and "emb_layers" not in k) else v) for k, v in sd.items()}
    oF = types.SimpleNamespace(**{k: getattr(F, k) for k in dir(F)})
    oE = torch.einsum
    class FF: pass
    ff = types.SimpleNamespace(**{k: getattr(F, k) for k in dir(F)})
    if round_a:
        ff.linear = lambda x, w, b=None: oF.linear(q(x) if x.shape[-1] != sd["time_embed.0.weight"].shape[0] and x.dim() > 2 else x, w, b)
        ff.conv2d = lambda x, w, b=None, stride=1, padding=0: oF.conv2d(q(x), w, b, stride=stride, padding=padding)
    O.F = ff
    if round_attn:
        class T:  # proxy for torch in oracle namespace
            def __getattr__(self, n): return getattr(torch, n)
            def einsum(self, eq, a, b): return oE(eq, q(a), q(b))
        O.torch = T()
    try:
        eps = O.unet_forward(sd, case["x"], case["t"], case["ctx"], num_heads=CFGS["unet"][case["cfg"]]["num_heads"])
    finally:
        O.F = F; O.torch = torch
    return rel_l2(eps, case["eps"])

cases = golden("unet.pt")
for idx in (0, 2):
    c = cases[idx]
    print(c["cfg"], tuple(c["x"].shape))
    print("  weights only     :", f"{run(c, True, False, False):.2e}")
    print("  activations only :", f"{run(c, False, True, False):.2e}")
    print("  attention only   :", f"{run(c, False, False, True):.2e}")
    print("  all              :", f"{run(c, True, True, True):.2e}")
