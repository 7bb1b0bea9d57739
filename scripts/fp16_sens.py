"""Which resolution levels / op kinds carry the fp16 operand-rounding error? (CPU emulation on the oracle)"""
import sys, torch, types
sys.path.insert(0, "."); sys.path.insert(0, "oracle"); sys.path.insert(0, "tests")
import torch.nn.functional as F
import ldm_oracle as O
from helpers import CFGS, golden, rel_l2, weights
q = lambda t: t.half().float()
case = golden("unet.pt")[2]   # sdv1 16x16
sd = weights("unet", "sdv1", case["seed"])
H = case["x"].shape[-1]

def run(pred):
    """pred(kind, ntok_or_hw, K) -> True to round both operands of that GEMM"""
    oF = F
    ff = types.SimpleNamespace(**{k: getattr(F, k) for k in dir(F)})
    def lin(x, w, b=None):
        if x.dim() > 2 and pred("linear", x.shape[1], w.shape[1], w.shape[0]):
            return oF.linear(q(x), q(w), b)
        return oF.linear(x, w, b)
    def conv(x, w, b=None, stride=1, padding=0):
        if pred("conv%d" % w.shape[-1], x.shape[-1] * x.shape[-2], w.shape[1] * w.shape[2] * w.shape[3], w.shape[0]):
            return oF.conv2d(q(x), q(w), b, stride=stride, padding=padding)
        return oF.conv2d(x, w, b, stride=stride, padding=padding)
    ff.linear, ff.conv2d = lin, conv
    O.F = ff
    try:
        eps = O.unet_forward(sd, case["x"], case["t"], case["ctx"])
    finally:
        O.F = F
    return rel_l2(eps, case["eps"])

print("all GEMMs (no attention rounding):", f"{run(lambda *a: True):.2e}")
for lvl in range(4):
    hw = (H >> lvl) ** 2
    e = run(lambda kind, n, K, N: n == hw)
    print(f"level {lvl} (hw={hw}) only: {e:.2e}   var share {e*e/ (1.17e-3**2):.2f}")
for kind in ("conv3", "conv1", "linear"):
    e = run(lambda k, n, K, N: k == kind)
    print(f"kind {kind} only: {e:.2e}")
e = run(lambda k, n, K, N: k == "conv3" and N == 4)
print(f"final out conv only: {e:.2e}")
e = run(lambda k, n, K, N: k == "linear" and n == 77)
print(f"context k/v only: {e:.2e}")
e = run(lambda k, n, K, N: k == "linear" and N >= 2560)
print(f"GEGLU proj only: {e:.2e}")
e = run(lambda k, n, K, N: k == "linear" and K >= 1280 and N <= 1280 and K == 4 * N)
print(f"FF out only: {e:.2e}")
