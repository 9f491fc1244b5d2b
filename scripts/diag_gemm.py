"""Bring-up diagnostics for the tcgen05 GEMMs: structured integer inputs that expose operand layout
bugs (which element landed where).  Prints a compact report; not a test."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from videovector_b200 import ops

torch.cuda.set_device(0)


def report(name, got, ref):
    got = got.double().cpu(); ref = ref.double().cpu()
    bad = (got - ref).abs() > 1e-3 * ref.abs().clamp_min(1.0)
    nbad = int(bad.sum())
    print("%-46s %s  max|err| %.3e  bad %d/%d" % (name, "OK " if nbad == 0 else "BAD", float((got - ref).abs().max()), nbad, got.numel()))
    if nbad:
        idx = bad.nonzero()[:6]
        for i, j in idx.tolist():
            print("     [%d,%d] got %.1f want %.1f" % (i, j, got[i, j], ref[i, j]))
        rows = bad.any(1).nonzero().flatten()
        cols = bad.any(0).nonzero().flatten()
        print("     bad rows: %d (first %s)  bad cols: %d (first %s)" % (len(rows), rows[:8].tolist(), len(cols), cols[:8].tolist()))


def run(prec):
    for (M, N, K) in [(128, 256, 64), (128, 256, 256), (256, 512, 128), (384, 256, 4096)]:
        X = (torch.arange(M, device="cuda").view(M, 1) % 13 + torch.arange(K, device="cuda").view(1, K) % 7).float()
        W = ((torch.arange(N, device="cuda").view(N, 1) % 5) - (torch.arange(K, device="cuda").view(1, K) % 3)).float()
        dZ = ((torch.arange(M, device="cuda").view(M, 1) % 3) + (torch.arange(N, device="cuda").view(1, N) % 11) - 4).float()
        Xo, Wo, dZo = ops.prepare_operand(X, prec), ops.prepare_operand(W, prec), ops.prepare_operand(dZ, prec)
        try:
            H, _ = ops.ip_forward(Xo, Wo, None, M, N, K, prec)
            torch.cuda.synchronize()
            report("%s fwd   M%d N%d K%d" % (prec, M, N, K), H, X.double() @ W.double().t())
        except Exception as e:
            print("fwd failed:", e)
        try:
            dW = ops.ip_wgrad(dZo, Xo, M, N, K, prec, nsplit=1)
            torch.cuda.synchronize()
            report("%s wgrad M%d N%d K%d" % (prec, M, N, K), dW, dZ.double().t() @ X.double())
        except Exception as e:
            print("wgrad failed:", e)
        try:
            dX = ops.ip_dgrad(dZo, Wo, M, N, K, prec)
            torch.cuda.synchronize()
            report("%s dgrad M%d N%d K%d" % (prec, M, N, K), dX, dZ.double() @ W.double())
        except Exception as e:
            print("dgrad failed:", e)


if __name__ == "__main__":
    for prec in (sys.argv[1:] or ["fp32_simt", "bf16", "tf32", "tf32x3"]):
        run(prec)
