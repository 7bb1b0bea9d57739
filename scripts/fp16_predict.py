"""Predicted eps error of the fp16x3 precision mode (CPU emulation on the oracle): every GEMM operand rounded to fp16
except the 1x1 convs and the final conv; attention operands rounded too."""
import sys, torch, types
sys.path.insert(0, "."); sys.path.insert(0, "oracle"); sys.path.insert(0, "tests")
import torch.nn.functional as F
import ldm_oracle as O
from helpers import golden, rel_l2, weights
q = lambda t: t.half().float()
for idx in (2, 3):
    case = golden("unet.pt")[idx]
    sd = weights("unet", "sdv1", case["seed"])
    ff = types.SimpleNamespace(**{k: getattr(F, k) for k in dir(F)})
    def lin(x, w, b=None):
        return F.linear(q(x), q(w), b) if x.dim() > 2 else F.linear(x, w, b)
    def conv(x, w, b=None, stride=1, padding=0):
        if w.shape[-1] == 1 or w.shape[0] == 4:
            return F.conv2d(x, w, b, stride=stride, padding=padding)
        return F.conv2d(q(x), q(w), b, stride=stride, padding=padding)
    ff.linear, ff.conv2d = lin, conv
    class T:
        def __getattr__(self, n): return getattr(torch, n)
        def einsum(self, eq, a, b): return torch.einsum(eq, q(a), q(b))
    O.F, O.torch = ff, T()
    try:
        eps = O.unet_forward(sd, case["x"], case["t"], case["ctx"])
    finally:
        O.F, O.torch = F, torch
    print(tuple(case["x"].shape), f"predicted rel-L2 {rel_l2(eps, case['eps']):.3e}")
