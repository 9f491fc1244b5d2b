"""How does the tf32x3 error grow with the accumulation chain length? (bias vs random walk)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from videovector_b200 import ops
torch.cuda.set_device(0)
g = torch.Generator(device="cuda").manual_seed(0)
M, N = 512, 512
for K in (256, 512, 1024, 2048, 4096, 8192, 16384):
    X = torch.relu(torch.randn(M, K, device="cuda", generator=g))
    W = torch.randn(N, K, device="cuda", generator=g) * 0.05
    ref = X.double() @ W.double().t()
    out = {}
    for prec in ("fp32_simt", "tf32x3"):
        H, _ = ops.ip_forward(ops.prepare_operand(X, prec), ops.prepare_operand(W, prec), None, M, N, K, prec)
        e = (H.double() - ref)
        relmax = float(e.abs().max() / ref.abs().max())
        rell2 = float(e.norm() / ref.norm())
        # signed bias: correlation of the error with the sign of the reference (negative => toward zero)
        bias = float((e * ref.sign()).mean() / ref.abs().mean())
        out[prec] = (relmax, rell2, bias)
    print("K=%6d  simt relmax %.2e l2 %.2e bias %+.2e | x3 relmax %.2e l2 %.2e bias %+.2e" % ((K,) + out["fp32_simt"] + out["tf32x3"]))
# all-positive operands: every partial sum grows monotonically -> worst case for truncation
for K in (1024, 4096):
    X = torch.rand(M, K, device="cuda", generator=g); W = torch.rand(N, K, device="cuda", generator=g)
    ref = X.double() @ W.double().t()
    H, _ = ops.ip_forward(ops.prepare_operand(X, "tf32x3"), ops.prepare_operand(W, "tf32x3"), None, M, N, K, "tf32x3")
    e = H.double() - ref
    print("positive K=%d: x3 relmax %.2e  mean signed rel err %+.2e" % (K, float(e.abs().max() / ref.abs().max()), float((e / ref).mean())))
